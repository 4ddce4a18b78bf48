/* vcprep.h — C ABI of libvcprep.so: the B200 (sm_100a) page-image prep path.
 *
 * Replaces, for the one hot path of the reference, the native routines Pillow/zlib/binascii run
 * when the reference does
 *     page_image.save(page_image_path)            backend/app/pipeline/pdf_extract.py:130
 *                                                 scripts/extract_pdf_with_gemini.py:152
 *                                                 scripts/extract_page_with_gemini.py:123
 *     model.generate_content([prompt, image])     backend/app/pipeline/pdf_extract.py:55   (image -> PNG blob [-> base64])
 * i.e. PIL convert('RGB') -> resize/thumbnail -> PNG (row filter + zlib stream + container) -> base64.
 * The reference has no plugin registry; the boundary is a plain function (SURVEY.md §8 b), bound from
 * Python with ctypes (vision_compression_project_b200/_native.py; INTEGRATION.md shows the call-site stub).
 *
 * Plain C: POD structs, raw pointers and sizes only.  Every function returns 0 or a negative VCP_E*;
 * the message of the last error on the calling thread is vcp_last_error().
 * A handle owns four lanes (CUDA stream + device arena + pinned staging each; host batches rotate over them so copies and
 * kernels overlap); calls on one handle are serialised by a mutex, different handles are independent (the reference calls the path
 * from 5 threads: pdf_extract.py:313-333).  Between vcp_batch_begin and vcp_batch_end the handle belongs to the streaming worker:
 * any other call on it returns VCP_EINVAL instead of blocking.
 */
#ifndef VCPREP_H
#define VCPREP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VCP_VERSION 101

#define VCP_OK            0
#define VCP_EINVAL       -1   /* bad argument / unsupported mode or size   (Python: ValueError)   */
#define VCP_ECUDA        -2   /* CUDA runtime error                        (Python: RuntimeError) */
#define VCP_ENOMEM       -3   /* device or host allocation failed          (Python: MemoryError)  */
#define VCP_ESIZE        -4   /* caller's output buffer too small          (Python: ValueError)   */

/* Pillow's Image.Resampling ids (PIL/Image.py) — same numbers so callers pass them through. */
#define VCP_LANCZOS  1
#define VCP_BILINEAR 2
#define VCP_BICUBIC  3
#define VCP_BOX      4
#define VCP_HAMMING  5

typedef struct vcp_handle vcp_handle;

/* One input page: row-major interleaved uint8 pixels, as Image.tobytes() / a PPM payload lays them out
 * (replaces the PIL.Image the reference holds at pdf_extract.py:129). */
typedef struct {
    const void* src;        /* host or device pointer, see vcp_opts.src_device                       */
    int32_t width, height;
    int32_t channels;       /* 1 = L, 2 = LA, 3 = RGB, 4 = RGBA                                      */
    int64_t row_stride;     /* bytes between rows; 0 = width*channels                                */
    int32_t dst_width;      /* target size of Image.resize; 0,0 = keep                               */
    int32_t dst_height;
    int32_t reduce_x;       /* Image.reduce factors applied before the resample (thumbnail's         */
    int32_t reduce_y;       /* reducing_gap step); 0 or 1 = none                                     */
    const void* const* row_ptrs;  /* optional (host sources only): `height` row pointers, row y = row_ptrs[y], for images whose
                                     rows do not lie at one stride — Pillow keeps images above 16 MB in several blocks
                                     (libImaging/Storage.c) and addresses them through exactly such a table; NULL = src + y*row_stride */
} vcp_page_desc;

typedef struct {
    int32_t out_channels;   /* 3 = convert('RGB'), 1 = convert('L')                                  */
    int32_t resample;       /* VCP_LANCZOS ...                                                        */
    int32_t compress_level; /* Pillow's compress_level: 0 = stored blocks; 1-3 fast, 4-6 default,      */
                            /* 7-9 best effort class of the GPU deflate (size/speed trade like zlib's)  */
    int32_t optimize;       /* 1 = Pillow optimize=True filter rule (Avg candidate included)          */
    int32_t want_b64;       /* 1 = also produce base64.b64encode(png)                                 */
    int32_t src_device;     /* 1 = page src pointers are device pointers (no H2D)                     */
    int32_t dst_device;     /* 1 = out_png / out_b64 are device pointers (no D2H of payloads)         */
    int32_t reserved;
} vcp_opts;

typedef struct {
    int32_t status;         /* VCP_OK or VCP_E* for this page (a bad page never fails the batch)      */
    int32_t width, height, channels;   /* of the encoded PNG                                          */
    uint64_t png_off, png_len;         /* byte range inside out_png                                   */
    uint64_t b64_off, b64_len;         /* byte range inside out_b64 (0,0 when want_b64 = 0)           */
    uint32_t adler32;                  /* of the filtered stream (zlib trailer)                       */
    uint32_t n_idat;
} vcp_page_result;

typedef struct {
    float ms_h2d, ms_convert, ms_resample, ms_filter, ms_lz, ms_huff, ms_assemble, ms_b64, ms_d2h, ms_total;
    uint64_t kernel_launches;          /* launches of this library's kernels in the last batch        */
    uint64_t in_bytes, filtered_bytes, png_bytes, b64_bytes;
    uint64_t arena_bytes;
} vcp_stats;

int         vcp_version(void);
const char* vcp_last_error(void);

int  vcp_init(int device, vcp_handle** out);
void vcp_destroy(vcp_handle* h);

/* Validate one page against the options without running anything: VCP_OK, or the VCP_E* that
 * vcp_prepare_batch would put in its status (message in vcp_last_error). */
int vcp_check_page(const vcp_page_desc* page, const vcp_opts* opts);

/* Worst-case output sizes for a batch, so the caller can size out_png / out_b64. */
int vcp_output_bound(const vcp_page_desc* pages, int n, const vcp_opts* opts,
                     uint64_t* png_bytes, uint64_t* b64_bytes);

/* The path: n pages -> n PNG byte strings (+ base64).  Synchronous: outputs are complete on return.
 * Replaces page_image.save(...) and the SDK's image->blob step for a whole batch in one launch set. */
int vcp_prepare_batch(vcp_handle* h, const vcp_page_desc* pages, int n, const vcp_opts* opts,
                      void* out_png, uint64_t png_cap, void* out_b64, uint64_t b64_cap,
                      vcp_page_result* results);

/* Streaming form of vcp_prepare_batch for bindings that want to consume results while later pages are still in flight:
 * vcp_batch_begin starts the same pipeline on a worker thread of the handle and returns at once; every vcp_batch_next blocks
 * until the next run of pages [first_page, last_page] has its bytes in out_png / out_b64 and its results[] final (returns 1),
 * returns 0 when the batch is finished, < 0 on error; vcp_batch_end joins the worker (always call it, from the thread that
 * called begin).  Buffers and the pages/results arrays must stay alive until vcp_batch_end. */
int vcp_batch_begin(vcp_handle* h, const vcp_page_desc* pages, int n, const vcp_opts* opts,
                    void* out_png, uint64_t png_cap, void* out_b64, uint64_t b64_cap, vcp_page_result* results);
int vcp_batch_next(vcp_handle* h, int* first_page, int* last_page);
int vcp_batch_end(vcp_handle* h);

/* ---- PNG decode (re-reading the images/page_###.png cache of backend/README.md:238-243, and checking
 * this library's own output at speed).  Replaces Image.open(png).load(): PngImagePlugin chunk parsing, zlib inflate and
 * libImaging/ZipDecode.c un-filtering.  8-bit gray / gray+alpha / RGB / RGBA, non-interlaced, any zlib stream.
 * Pixels come out interleaved, W*H*C per page, pages 256-byte aligned inside out_pixels. */
typedef struct {
    int32_t status;                    /* VCP_OK or VCP_E* for this PNG                                  */
    int32_t width, height, channels;
    uint64_t pix_off, pix_len;         /* byte range inside out_pixels                                   */
} vcp_decode_result;
/* Accepts and rejects what Image.open(png).load() does (Pillow 12 / zlib 1.3): chunks in front of the first IDAT need a correct CRC-32
 * (IDAT CRCs are skipped like Pillow skips them), the zlib header, every block header and code must be valid (incomplete codes are
 * errors as in inftrees.c), the stream must deliver every row, and the Adler-32 of the inflated stream — computed on the device — must
 * match the trailer whenever zlib's inflate() would have met it in the call that produced the last row.  status = VCP_EINVAL otherwise. */
int vcp_png_decode_batch(vcp_handle* h, const void* const* pngs /* host pointers */, const uint64_t* png_lens, int n,
                         void* out_pixels, uint64_t out_cap, int dst_device, vcp_decode_result* results);

int vcp_get_stats(vcp_handle* h, vcp_stats* out);

/* Host-side helper for language bindings: copy n byte ranges src_base+offs[i] .. +lens[i] into dsts[i] on `threads`
 * host threads (the Python binding fills freshly allocated bytes objects with it while the GIL is released). */
int vcp_host_scatter(const void* src_base, const uint64_t* offs, const uint64_t* lens, void* const* dsts, int n, int threads);

/* ---- stage-level entry points (device pointers; used by the parity tests, one image at a time) ---- */

/* Image.convert: Convert.c  (L/LA/RGB/RGBA -> RGB, RGB/RGBA/LA -> L). */
int vcp_convert(vcp_handle* h, const void* d_src, int width, int height, int src_channels, int64_t row_stride,
                void* d_dst, int dst_channels);
/* Resample.c precompute_coeffs + normalize_coeffs_8bpc, on the host (double math). bounds = out*2 int32
 * (xmin, n); kk = out*ksize int32 Q22.  Pass kk = NULL to query ksize only. */
int vcp_resample_coeffs(int in_size, int out_size, int filter, float box0, float box1,
                        int32_t* bounds, int32_t* kk, int* ksize);
/* Image.resize (8-bit, horizontal pass then vertical). d_dst is out_w*out_h*channels. */
int vcp_resample(vcp_handle* h, const void* d_src, int width, int height, int channels,
                 void* d_dst, int out_width, int out_height, int filter);
/* Image.reduce: Reduce.c. d_dst is ceil(w/fx)*ceil(h/fy)*channels. */
int vcp_reduce(vcp_handle* h, const void* d_src, int width, int height, int channels,
               void* d_dst, int fx, int fy);
/* ZipEncode.c adaptive filter: d_dst receives height*(1+width*channels) bytes. */
int vcp_png_filter(vcp_handle* h, const void* d_pix, int width, int height, int channels, int optimize,
                   void* d_dst, uint32_t* adler32_out);
/* zlib stream (header + deflate blocks + Adler-32) of one byte stream. bpp = pixel stride hint. */
int vcp_deflate(vcp_handle* h, const void* d_stream, uint64_t len, int bpp, int level,
                void* d_out, uint64_t cap, uint64_t* out_len);
/* LZ77 token stream of the deflate (test hook; compared against tests/model/deflate_model.c).
 * The stream is cut into deflate blocks of 512 KiB and those into sub-chunks of vcp_lz_sub_bytes() (one warp each).
 * d_tokens: len uint32, the tokens of a sub-chunk lie compact from its first byte position; sub_ntok_host: one uint32 per
 * sub-chunk; sub_hist_host: 316 per sub-chunk or NULL. */
int vcp_lz_sub_bytes(void);
int vcp_lz_tokens(vcp_handle* h, const void* d_stream, uint64_t len, int bpp,
                  uint32_t* d_tokens, uint32_t* sub_ntok_host, uint32_t* sub_hist_host /* nsub*316 or NULL */);
int vcp_adler32(vcp_handle* h, const void* d_data, uint64_t len, uint32_t* out);
int vcp_crc32(vcp_handle* h, const void* d_data, uint64_t len, uint32_t* out);
/* base64.b64encode: d_dst receives 4*ceil(len/3) bytes. */
int vcp_base64(vcp_handle* h, const void* d_src, uint64_t len, void* d_dst);

#ifdef __cplusplus
}
#endif
#endif /* VCPREP_H */
