// vcp_internal.cuh — device-side descriptors and kernel launchers shared by the .cu files of libvcprep.
// Everything here is internal; the public surface is include/vcprep.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vcp {

#ifndef VCP_UF_GROUP
#define VCP_UF_GROUP 4
#endif
#ifndef VCP_SUB_BYTES
#define VCP_SUB_BYTES 16384
#endif
#ifndef VCP_PRIME_BYTES
#define VCP_PRIME_BYTES 16384
#endif
// Measured on B200 (gpurun_out/r2_variants*.log -> profiles/r02_lz_granularity.md): the LZ stage of a launch set ends with its heaviest
// sub-chunk (one warp walking dense glyph rows), and hashing the history in front of a sub-chunk costs as much per byte as a quarter
// of the parse.  16 KiB sub-chunks primed with 16 KiB: same kernel time as 32/32 on a full batch, half the latency of a small one
// (single page 2.85 -> 2.0 ms, the reference's 5-thread pattern 1600 -> 2100 pages/s), +1.3 % PNG size on text pages.
#ifndef VCP_SPEC_BITS
#define VCP_SPEC_BITS 32768        // decode: at most one speculative parse start per 4 KiB of compressed stream (png_decode.cu k_infl_spec)
#endif
constexpr int kSubBytes   = VCP_SUB_BYTES;   // LZ sub-chunk: one warp; table entries are u16 positions relative to (start - 32 KiB)
constexpr int kPrimeBytes = VCP_PRIME_BYTES; // bytes in front of a sub-chunk that are hashed into its tables before it starts (<= 32 KiB, multiple of 512)
constexpr int kBlockBytes = 512 * 1024;   // deflate block = one IDAT chunk = 16 sub-chunks
constexpr int kGroupSubs  = 1;            // consecutive sub-chunks of a page one warp handles with one pair of hash tables
                                          // (2 halves the priming work but measured slower on B200: fewer, longer work items -> ragged tail)
constexpr int kMaxDist    = 32768;
constexpr int kMaxMatch   = 258;
constexpr int kNumLL      = 286;
constexpr int kNumD       = 30;
constexpr int kHistSize   = kNumLL + kNumD;   // 316
constexpr int kCodeStride = 320;          // per-block code table stride (u16 codes / u8 lengths)
constexpr int kHdrBytes   = 768;          // per-block dynamic-header scratch (<= 17 + 57 + 316*14 bits)
constexpr int kMaxDevices = 64;           // per-device caches of kernel attributes (a process may drive every GPU of the box)
constexpr int kStreamPad  = 256;          // bytes of addressable slack before and after every page stream

// One page of a batch (lives in device memory, built on the host per plan).
struct PageD {
    const uint8_t* src;      // source pixels, sc channels
    int64_t src_stride;
    int32_t sw, sh, sc;      // source geometry
    int32_t c;               // output channels (1 or 3 [or 2/4 when kept])
    int32_t pc;              // channels the pixel stages work in; pc = 1 with c = 3 means gray kept single-channel until the
                             // PNG filter stages it (L/LA -> RGB replicates, so resampling one channel is exact and 3x cheaper)
    int32_t fx, fy;          // reduce factors (1 = none)
    int32_t rw, rh;          // geometry after convert+reduce
    int32_t w, h;            // final geometry
    uint8_t* conv;           // sw*sh*c   convert output (nullptr: stage skipped)
    const uint8_t* rdin; int64_t rdin_stride;   // reduce input
    uint8_t* red;            // rw*rh*c   reduce output (nullptr: stage skipped)
    const uint8_t* hin; int64_t hin_stride;     // horizontal-pass input (rw x rh)
    uint8_t* tmp;            // w*rh*c    horizontal-pass output (nullptr: stage skipped)
    int64_t tmp_stride;      // bytes between rows of tmp (padded to 16 inside the arena)
    const uint8_t* vin; int64_t vin_stride;     // vertical-pass input (w x rh)
    uint8_t* vout;           // w*h*c     vertical-pass output (nullptr: stage skipped)
    int64_t vout_stride;     // bytes between rows of vout (padded to 16 inside the arena)
    const uint8_t* pix;      // w*h*c     final pixels (aliases src/conv/red/tmp/vout)
    int64_t pix_stride;      // bytes between rows of pix
    const int32_t* hb; const int32_t* hk; int32_t hks;   // horizontal bounds (xmin,n)*w, coeffs [k][w] (transposed), ksize
    const int32_t* vb; const int32_t* vk; int32_t vks;   // vertical
    uint8_t* filt;           // filtered stream: h * (1 + w*c) bytes (256-byte aligned, kStreamPad slack either side)
    int64_t filt_len;
    int32_t row0;            // first entry of this page in the per-row Adler partial array
    int32_t blk0, nblk;      // deflate blocks of this page
    int32_t color_type;      // PNG colour type of the output (0 L, 2 RGB, 4 LA, 6 RGBA)
    int32_t status;
};

// One deflate block = one IDAT chunk.
struct BlockD {
    int32_t page;
    int32_t first, last;     // first / last block of its page
    int64_t start, len;      // byte range inside the page's filtered stream
    int32_t sub0, nsub;      // LZ sub-chunks of this block (global numbering)
};

struct BatchD {
    const PageD* pages; int32_t npages;
    const BlockD* blocks; int32_t nblocks;
    const uint32_t* sub2blk; int32_t nsub;
    const uint32_t* item2sub; int32_t nitems;   // LZ work items: first sub-chunk of each group of <= kGroupSubs (never across pages)
    const uint8_t* filt_base;    // base of the filtered buffer (token index = filt ptr - filt_base)
    uint32_t* tokens;        // u32 per filtered byte position; tokens of a sub-chunk are compact from its first position
    uint32_t* sub_ntok;      // tokens per sub-chunk
    uint32_t* sub_hist;      // nsub * 316 histogram (without EOB)
    uint32_t* row_adler;     // per row partials, see png_filter.cu
    uint32_t* page_adler;    // per page
    // per deflate block, written by k_huff_build
    uint16_t* blk_code;      // nblocks * kCodeStride: bit-reversed canonical codes, [0,286) lit/len, [286,316) dist
    uint8_t*  blk_clen;      // nblocks * kCodeStride: code lengths
    uint8_t*  blk_hdr;       // nblocks * kHdrBytes: BFINAL/BTYPE + dynamic header bits
    uint32_t* blk_hdr_bits;
    uint64_t* blk_eob_bit;   // bit offset of the EOB code inside the block body
    uint64_t* blk_body_bits; // total bits of the body (incl. EOB and, if not last, the 3-bit empty stored header)
    uint32_t* blk_stored;    // 1 = stored fallback
    uint32_t* blk_len;       // payload bytes (zlib header + body + sync marker / Adler-32)
    uint64_t* sub_bitoff;    // per sub-chunk: bit offset of its first token inside the block body
    // layout (k_layout)
    uint64_t* blk_dst;       // byte offset in png of the block's payload (after the 8-byte chunk header when framed)
    uint8_t* png; uint64_t png_cap;
    uint64_t* png_off; uint64_t* png_len;   // per page (offsets 16-byte aligned)
    uint8_t* b64; uint64_t b64_cap;
    uint64_t* b64_off; uint64_t* b64_len;
    uint64_t* totals;        // [0] = png bytes used, [1] = b64 bytes used
    uint32_t* err;           // [0] != 0: capacity exceeded
    uint32_t* counters;      // zeroed per launch set: [0] = next LZ sub-chunk (work queue of k_lz), [64..192) = histogram and cursors of the work-item sort
    uint8_t* row_busy;       // per row: 1 = the PNG filter found content (0 = identical to the row above), nullptr when unknown
    uint32_t* lz_order;      // LZ work items sorted by estimated cost, heaviest first (nullptr: stream order)
    int32_t framed;          // 1 = PNG container (sig/IHDR/IDAT/IEND); 0 = bare zlib stream (vcp_deflate)
    int32_t level;           // 0 = stored only
    int32_t want_b64;
};

// One PNG being decoded (png_decode.cu).
struct DecPageD {
    const uint8_t* z; unsigned long long zlen;      // concatenated IDAT payloads (one zlib stream)
    uint8_t* filt; unsigned long long filt_len;     // h * (1 + w*c)
    uint8_t* pix;                                   // w*h*c
    uint16_t* sym;                                  // filt_len symbolic bytes (see k_infl_exec)
    int32_t w, h, c;
    int32_t status;                                 // 0 or a negative inflate / un-filter error
    int32_t seg0, seg_cap, nseg;                    // its parse units in the DecSegD array (nseg written by k_infl_sort)
    int32_t cand0, n_idat;                          // its range of candidate start bits; the first n_idat are the IDAT starts (host)
    int32_t surv0, surv_cap;                        // its range of k_infl_scan1 survivors
    uint32_t nsurv, ncand;                          // device counters (ncand starts at n_idat)
    int32_t slot0, page_iv;                         // first checkpoint slot; slots per parse unit (filt_len / 64 Ki + 2)
    int32_t iv0, iv_cap, niv;                       // its intervals in the DecIvD array (niv written by k_infl_plan)
    int32_t band0;                                  // first 32-row band of this page in the un-filter's band numbering
    int32_t chunk0;                                 // first 32 KiB chunk of this page in the resolve / Adler work list
    int32_t spec0, nspec;                           // its range of speculative start points (one per spec_bits of stream, k_infl_spec)
    uint32_t spec_bits, pad_spec;                   // distance of those points in the stream, bits (host: 8192 .. VCP_SPEC_BITS by batch size and stream length)
    // ---- what Pillow's decoder (zlib inflate driven row by row, ZipDecode.c) would have seen at the end of the image; the host
    //      turns these into accept / reject exactly as Image.open(png).load() does (api.cu: decode_verdict)
    unsigned long long valid_len;                   // inflated bytes that exist: filt_len, or less when the final block ended early
    unsigned long long done_bit;                    // bit position behind the token that produced the last byte of the image (0: it ended inside a token)
    unsigned long long end_bit;                     // bit position behind the final block when nothing but block ends / headers lie between (0: not reached)
    unsigned long long post_err_bit;                // where a malformed element sits between done_bit and the first thing that needs output space
    int32_t post_err;                               // its error code (0: none)
    uint32_t adler;                                 // Adler-32 of the valid_len inflated bytes (k_dec_adler_fin)
};

// A place where a parse of a page's deflate stream may begin: an IDAT start or a block header found by the scan (k_infl_probe).
struct DecSegD {
    uint32_t page, pad;
    unsigned long long start_bit;   // bit position inside the page's zlib stream
    uint32_t olen, opos;     // bytes its parse produced (k_infl_probe) and where they start in the filtered stream (k_infl_plan)
    int32_t ok;              // 1 = the parse ended on a block boundary that is another parse unit's start (or the end of the stream); < 0 error
    int32_t next;            // the parse unit (index within the page) in front of which it ended
    int32_t fin;             // it met the final block
    uint32_t niv;            // intervals it wrote
    uint32_t iv0, iv_cap;    // its private range of checkpoint slots
    unsigned long long end_bit;     // fin: bit position behind the final block (the Adler-32 follows at the next byte boundary)
    unsigned long long hdr_bit;     // header of the deflate block the unit starts in (== start_bit unless the unit starts inside a block, k_infl_spec)
};

// A stretch of tokens between two checkpoints of a parse: the unit of k_infl_exec.
struct DecIvD {
    unsigned long long hdr_bit;    // bit position (in the page's zlib stream) of the header of the deflate block it starts in
    unsigned long long start_bit;  // bit position of its first token
    uint32_t out, len;             // output range (relative to the parse unit in the probe's slots, absolute after k_infl_plan)
    uint32_t seg;
    uint32_t last;                 // 1 = the interval that ends the image: its last token may reach past the end (clipped), and what follows is inspected
};

struct DecBatchD {
    DecPageD* pages; int32_t npages;
    DecSegD* segs; int32_t seg_total;                                         // parse units, one range of seg_cap per page
    unsigned long long* cand_bits;                                            // candidate start bits, same ranges
    unsigned long long* cand_hdr;                                             // per candidate: header bit of its block (0: the candidate IS a block start)
    int32_t spec_total;                                                       // speculative start points over all pages
    uint32_t* surv; int32_t surv_total;                                       // k_infl_scan1 survivors, one range per page
    const uint32_t* scan_page; const uint32_t* scan_bit; int32_t nscan;       // scan work list: 8 Ki bit positions each
    DecIvD* slots;                                                            // checkpoint slots, one private range per parse unit
    DecIvD* ivs; int32_t iv_total;                                            // intervals in stream order, one range per page
    const uint32_t* chunk_page; const uint32_t* chunk_pos; int32_t nchunks;   // resolve / Adler work list: 32 Ki positions each
    uint32_t* chunk_adler;                                                    // per chunk: (sum of bytes, position-weighted sum) mod 65521
    uint32_t* band_flag; int32_t nbands; int32_t ngroups;                     // un-filter progress per band (zeroed per launch); CTAs = groups of VCP_UF_GROUP bands
    const uint32_t* band_page; const uint32_t* band_idx;                      // un-filter work list in ticket order (group-major over pages): page, first band
    uint32_t* counters;                                                       // [0] un-filter CTA ticket (zeroed per launch)
    int32_t dbg_nowait;                                                       // timing experiments only: bands do not wait (wrong pixels)
    int32_t no_scan;                                                          // VCP_DECODE_NO_SCAN: parse units are the IDAT starts only
};

// ---- TMA 1-D bulk copies (cp.async.bulk, global -> shared, completion on an mbarrier).  Used to stage byte rows whose global
// addresses share no word phase (3-byte pixels: a 2550-pixel RGB row is 7650 bytes): the copy engine needs only 16-byte alignment,
// so a row is fetched from its address rounded down to 16 and lands with a per-row byte offset (address & 15) in shared memory.
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");     // make the init visible to the async proxy
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {   // 16-byte aligned, size % 16 == 0
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}"
                 :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
#endif

// ---- launchers (each returns the number of kernels it launched) ----
int launch_inflate(const DecBatchD& b, cudaStream_t st);
int launch_unfilter(const DecBatchD& b, cudaStream_t st);
int launch_convert(const PageD* d_pages, int npages, int max_rows, int max_w, cudaStream_t st);
int launch_reduce(const PageD* d_pages, int npages, int max_rh, int max_rw, cudaStream_t st);
int launch_resample_h(const PageD* d_pages, int npages, int max_rh, int max_w, cudaStream_t st);
int launch_resample_v(const PageD* d_pages, int npages, int max_h, int max_wc, cudaStream_t st);
int launch_png_filter(const PageD* d_pages, int npages, int max_h, int max_rowbytes, int optimize,
                      uint32_t* row_adler, uint8_t* row_busy, cudaStream_t st);
int launch_lz_order(const BatchD& b, cudaStream_t st);
int launch_adler_combine(const PageD* d_pages, int npages, const uint32_t* row_adler, uint32_t* page_adler, cudaStream_t st);
int launch_lz(const BatchD& b, cudaStream_t st);
int launch_huff_build(const BatchD& b, cudaStream_t st);
int launch_layout(const BatchD& b, cudaStream_t st);
int launch_payload_init(const BatchD& b, cudaStream_t st);
int launch_huff_emit(const BatchD& b, cudaStream_t st);
int launch_png_finish(const BatchD& b, cudaStream_t st);
int launch_base64_pages(const BatchD& b, cudaStream_t st);
int launch_base64_flat(const uint8_t* src, uint64_t len, uint8_t* dst, cudaStream_t st);
int launch_adler_flat(const uint8_t* data, uint64_t len, uint32_t* scratch, uint32_t* out, cudaStream_t st);
int launch_crc_flat(const uint8_t* data, uint64_t len, uint32_t* out, cudaStream_t st);

// host helper shared by api.cu and tests
int resample_coeffs_host(int in_size, int out_size, int filter, float box0, float box1,
                         int32_t* bounds, int32_t* kk, int* ksize);

}  // namespace vcp
