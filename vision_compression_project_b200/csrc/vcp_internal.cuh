// vcp_internal.cuh — device-side descriptors and kernel launchers shared by the .cu files of libvcprep.
// Everything here is internal; the public surface is include/vcprep.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vcp {

constexpr int kSubBytes   = 32768;        // LZ sub-chunk: one warp, one 64 Ki-position window of u16 table entries
constexpr int kMaxDist    = 32768;
constexpr int kMaxMatch   = 258;
constexpr int kNumLL      = 286;
constexpr int kNumD       = 30;
constexpr int kHistSize   = kNumLL + kNumD;   // 316
constexpr int kStreamPad  = 256;          // bytes of addressable slack before and after every page stream

// One page of a batch (lives in device memory, built on the host per plan).
struct PageD {
    const uint8_t* src;      // source pixels, sc channels
    int64_t src_stride;
    int32_t sw, sh, sc;      // source geometry
    int32_t c;               // output channels (1 or 3 [or 2/4 when kept])
    int32_t fx, fy;          // reduce factors (1 = none)
    int32_t rw, rh;          // geometry after convert+reduce
    int32_t w, h;            // final geometry
    uint8_t* conv;           // sw*sh*c   convert output (nullptr: stage skipped)
    const uint8_t* rdin; int64_t rdin_stride;   // reduce input
    uint8_t* red;            // rw*rh*c   reduce output (nullptr: stage skipped)
    const uint8_t* hin; int64_t hin_stride;     // horizontal-pass input (rw x rh)
    uint8_t* tmp;            // w*rh*c    horizontal-pass output (nullptr: stage skipped)
    const uint8_t* vin; int64_t vin_stride;     // vertical-pass input (w x rh)
    uint8_t* vout;           // w*h*c     vertical-pass output (nullptr: stage skipped)
    const uint8_t* pix;      // w*h*c     final pixels (aliases src/conv/red/tmp/vout)
    int64_t pix_stride;      // bytes between rows of pix
    const int32_t* hb; const int32_t* hk; int32_t hks;   // horizontal bounds (xmin,n)*w, coeffs [k][w] (transposed), ksize
    const int32_t* vb; const int32_t* vk; int32_t vks;   // vertical
    uint8_t* filt;           // filtered stream: h * (1 + w*c) bytes
    int64_t filt_len;
    int32_t row0;            // first entry of this page in the per-row Adler partial array
    int32_t blk0, nblk;      // deflate blocks of this page
    int32_t status;
};

// One deflate block = one IDAT chunk.
struct BlockD {
    int32_t page;
    int32_t first, last;     // first / last block of its page
    int64_t start, len;      // byte range inside the page's filtered stream
    int32_t sub0, nsub;      // LZ sub-chunks of this block (global numbering)
    int64_t slot_off;        // offset of this block's output slot in the slot buffer
    int64_t slot_cap;
};

struct BatchD {
    const PageD* pages; int32_t npages;
    const BlockD* blocks; int32_t nblocks;
    const uint32_t* sub2blk; int32_t nsub;
    uint32_t* tokens;        // u32 per filtered byte position (token j of sub-chunk at stream offset o -> tokens[tok_base(o) + j])
    uint32_t* sub_ntok;      // tokens per sub-chunk
    uint32_t* sub_hist;      // nsub * 316 histogram (without EOB)
    uint32_t* row_adler;     // per row: (sum & 0xFFFF) | (weighted << 16) ... see png_filter.cu
    uint32_t* page_adler;    // per page
    uint8_t* slots;          // per-block compressed payloads
    uint32_t* blk_len;       // payload bytes per block
    uint32_t* blk_crc;       // CRC-32 of "IDAT"+payload per block
    uint8_t* png;            // assembled PNGs
    uint64_t* png_off; uint64_t* png_len;   // per page
    uint8_t* b64;
    uint64_t* b64_off; uint64_t* b64_len;
    int64_t tok_base_of_filt0;   // unused (tokens are indexed by filtered-buffer byte offset)
    const uint8_t* filt_base;    // base of the filtered buffer (token index = filt ptr - filt_base)
};

// ---- launchers (each returns the number of kernels it launched) ----
int launch_convert(const PageD* d_pages, int npages, int max_rows, int max_rowbytes, cudaStream_t st);
int launch_reduce(const PageD* d_pages, int npages, int max_rh, int max_rw, cudaStream_t st);
int launch_resample_h(const PageD* d_pages, int npages, int max_rh, int max_w, cudaStream_t st);
int launch_resample_v(const PageD* d_pages, int npages, int max_h, int max_wc, cudaStream_t st);
int launch_png_filter(const PageD* d_pages, int npages, int max_h, int max_rowbytes, int optimize,
                      uint32_t* row_adler, cudaStream_t st);
int launch_adler_combine(const PageD* d_pages, int npages, const uint32_t* row_adler, uint32_t* page_adler, cudaStream_t st);
int launch_lz(const BatchD& b, int bpp_hint_unused, cudaStream_t st);
int launch_huff(const BatchD& b, int level, cudaStream_t st);
int launch_assemble(const BatchD& b, uint64_t png_cap, uint32_t* d_err, cudaStream_t st);
int launch_base64(const BatchD& b, uint64_t b64_cap, uint32_t* d_err, cudaStream_t st);
int launch_base64_flat(const uint8_t* src, uint64_t len, uint8_t* dst, cudaStream_t st);
int launch_adler_flat(const uint8_t* data, uint64_t len, uint32_t* scratch, uint32_t* out, cudaStream_t st);
int launch_crc_flat(const uint8_t* data, uint64_t len, uint32_t* out, cudaStream_t st);

// host helper shared by api.cu and tests
int resample_coeffs_host(int in_size, int out_size, int filter, float box0, float box1,
                         int32_t* bounds, int32_t* kk, int* ksize);

}  // namespace vcp
