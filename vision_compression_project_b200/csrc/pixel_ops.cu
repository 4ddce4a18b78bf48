// pixel_ops.cu — mode conversion, integer box reduce and Pillow-exact fixed-point resampling.
//
// Replaces (bit-exactly) what Pillow runs for the north-star stages in front of the PNG encoder:
//   Image.convert  -> libImaging/Convert.c      (PIL/Image.py:1018)
//   Image.reduce   -> libImaging/Reduce.c       (PIL/Image.py:2440)
//   Image.resize   -> libImaging/Resample.c     (PIL/Image.py:2328): horizontal pass into a uint8
//                     temporary, then vertical pass; Q22 int32 coefficients computed on the host in
//                     double precision (resample_coeffs_host) exactly as precompute_coeffs +
//                     normalize_coeffs_8bpc do, so no floating point ever touches a pixel on the device.
// All integer work, HBM/L2 bound: threads walk the contiguous (byte) dimension so warps read and write
// whole sectors; coefficient tables are stored transposed ([tap][out index]) so a warp's tap loads coalesce.
#include "vcp_internal.cuh"
#include <math.h>
#include <vector>

namespace vcp {

// ------------------------------------------------------------------------------------------ convert
// Image.convert (Convert.c): RGB(A) -> RGB drops the 4th byte, L/LA -> RGB replicates, RGB(A)/LA -> L takes Pillow's integer luma
// (R*19595 + G*38470 + B*7471 + 0x8000) >> 16.  A thread owns 4 consecutive pixels: 4-byte pixels are read as words (one 128-bit
// load when the row allows), 3-byte pixels as three words, and the 12 or 4 output bytes leave as words when the output row is word
// aligned (page rows of 4k pixels always are); everything else goes byte by byte.
constexpr int kConvRows = 8;
__device__ __forceinline__ uint32_t luma8(uint32_t r, uint32_t g, uint32_t b) { return (r * 19595u + g * 38470u + b * 7471u + 0x8000u) >> 16; }

__global__ void __launch_bounds__(256) k_convert(const PageD* __restrict__ pages) {
    const PageD& P = pages[blockIdx.z];
    if (!P.conv) return;
    const int x4 = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (x4 >= P.sw) return;
  for (int y = blockIdx.y * kConvRows; y < min(P.sh, (int)(blockIdx.y + 1) * kConvRows); y++) {     // several rows per CTA: fewer, longer-lived blocks
    const int sc = P.sc, c = P.pc;
    const uint8_t* __restrict__ srow = P.src + (int64_t)y * P.src_stride;
    uint8_t* __restrict__ drow = P.conv + (int64_t)y * P.sw * c;
    const bool full = x4 + 4 <= P.sw;
    const bool src_w = (((uintptr_t)srow) & 3) == 0, dst_w = (((uintptr_t)drow) & 3) == 0;
    if (full && sc == 4 && src_w) {
        uint32_t px[4];
        if ((((uintptr_t)srow) & 15) == 0) { const uint4 q = __ldg(reinterpret_cast<const uint4*>(srow) + (x4 >> 2)); px[0] = q.x; px[1] = q.y; px[2] = q.z; px[3] = q.w; }
        else { const uint32_t* g = reinterpret_cast<const uint32_t*>(srow) + x4; px[0] = __ldg(g); px[1] = __ldg(g + 1); px[2] = __ldg(g + 2); px[3] = __ldg(g + 3); }
        if (c == 3) {
            const uint32_t o0 = __byte_perm(px[0], px[1], 0x4210), o1 = __byte_perm(px[1], px[2], 0x5421), o2 = __byte_perm(px[2], px[3], 0x6542);
            if (dst_w) { uint32_t* d = reinterpret_cast<uint32_t*>(drow + (int64_t)x4 * 3); d[0] = o0; d[1] = o1; d[2] = o2; }
            else { uint8_t* d = drow + (int64_t)x4 * 3; const uint32_t o[3] = {o0, o1, o2};
#pragma unroll
                for (int k = 0; k < 12; k++) d[k] = (uint8_t)(o[k >> 2] >> (8 * (k & 3))); }
        } else {
            uint32_t o = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) o |= luma8(px[k] & 255u, (px[k] >> 8) & 255u, (px[k] >> 16) & 255u) << (8 * k);
            if (dst_w) reinterpret_cast<uint32_t*>(drow)[x4 >> 2] = o;
            else for (int k = 0; k < 4; k++) drow[x4 + k] = (uint8_t)(o >> (8 * k));
        }
        continue;
    }
    if (full && sc == 3 && c == 1 && src_w) {
        const uint32_t* g = reinterpret_cast<const uint32_t*>(srow + (int64_t)x4 * 3);
        const uint32_t a = __ldg(g), b = __ldg(g + 1), d2 = __ldg(g + 2);      // R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3
        const uint32_t o = luma8(a & 255u, (a >> 8) & 255u, (a >> 16) & 255u) | (luma8(a >> 24, b & 255u, (b >> 8) & 255u) << 8) |
                           (luma8((b >> 16) & 255u, b >> 24, d2 & 255u) << 16) | (luma8((d2 >> 8) & 255u, (d2 >> 16) & 255u, d2 >> 24) << 24);
        if (dst_w) reinterpret_cast<uint32_t*>(drow)[x4 >> 2] = o;
        else for (int k = 0; k < 4; k++) drow[x4 + k] = (uint8_t)(o >> (8 * k));
        continue;
    }
    for (int x = x4; x < min(P.sw, x4 + 4); x++) {
        const uint8_t* s = srow + (int64_t)x * sc;
        uint8_t* d = drow + (int64_t)x * c;
        if (c == 3) {
            if (sc <= 2) { uint8_t v = __ldg(s); d[0] = v; d[1] = v; d[2] = v; }            // L, LA -> RGB (alpha dropped)
            else { d[0] = __ldg(s); d[1] = __ldg(s + 1); d[2] = __ldg(s + 2); }             // RGB(A) -> RGB
        } else {                                                                            // -> L
            if (sc <= 2) d[0] = __ldg(s);
            else d[0] = (uint8_t)luma8(__ldg(s), __ldg(s + 1), __ldg(s + 2));
        }
    }
  }
}

int launch_convert(const PageD* d_pages, int npages, int max_rows, int max_w, cudaStream_t st) {
    if (npages == 0 || max_rows == 0) return 0;
    dim3 grid(((max_w + 3) / 4 + 255) / 256, (max_rows + kConvRows - 1) / kConvRows, npages);
    k_convert<<<grid, 256, 0, st>>>(d_pages);
    return 1;
}

// ------------------------------------------------------------------------------------------ reduce
// Image.reduce (Reduce.c): out = ((sum + n/2) * floor(2^24 / n)) >> 24 over the fx x fy cell, n = pixels actually inside the image.
// A thread owns 4 consecutive output pixels: their cells are 4*fx*c contiguous bytes of every input row, read as aligned words when
// the rows are word aligned (fx*c words per row instead of 4*fx*c byte loads); cells cut by the right edge, rows that are not word
// aligned and factors without an instantiation take the byte path.
__device__ __forceinline__ void reduce_pixel_bytes(const PageD& P, int ox, int oy) {
    const int c = P.pc, fx = P.fx, fy = P.fy;
    const int y0 = oy * fy, y1 = min(y0 + fy, P.sh);
    const int x0 = ox * fx, x1 = min(x0 + fx, P.sw);
    const uint32_t n = (uint32_t)(y1 - y0) * (uint32_t)(x1 - x0);
    const uint32_t mult = (1u << 24) / n;
    for (int ch = 0; ch < c; ch++) {
        uint32_t s = 0;
        for (int y = y0; y < y1; y++) {
            const uint8_t* r = P.rdin + (int64_t)y * P.rdin_stride + ch;
            for (int x = x0; x < x1; x++) s += __ldg(r + (int64_t)x * c);
        }
        P.red[((int64_t)oy * P.rw + ox) * c + ch] = (uint8_t)(((s + n / 2) * mult) >> 24);
    }
}

template <int FX, int C>
__device__ __forceinline__ bool reduce_quad_words(const PageD& P, int ox4, int oy) {
    // all four cells inside the image, rows word aligned?
    if ((ox4 + 4) * FX > P.sw || ((((uintptr_t)P.rdin) | (uintptr_t)P.rdin_stride) & 3) != 0) return false;
    const int fy = P.fy;
    const int y0 = oy * fy, y1 = min(y0 + fy, P.sh);
    constexpr int W = FX * C;                                   // words per input row for 4 output pixels
    uint32_t s[4][C];
#pragma unroll
    for (int p = 0; p < 4; p++)
#pragma unroll
        for (int ch = 0; ch < C; ch++) s[p][ch] = 0;
    for (int y = y0; y < y1; y++) {
        const uint32_t* __restrict__ rw = reinterpret_cast<const uint32_t*>(P.rdin + (int64_t)y * P.rdin_stride + (int64_t)ox4 * W);
        uint32_t w[W];
#pragma unroll
        for (int k = 0; k < W; k++) w[k] = __ldg(rw + k);
#pragma unroll
        for (int bi = 0; bi < 4 * W; bi++)                      // byte bi of the run: pixel bi / (FX*C) of the quad, channel bi % C
            s[bi / W][bi % C] += __byte_perm(w[bi >> 2], 0u, 0x4440u | (uint32_t)(bi & 3));
    }
    const uint32_t n = (uint32_t)(y1 - y0) * FX;
    const uint32_t mult = (1u << 24) / n;
    uint8_t* o = P.red + ((int64_t)oy * P.rw + ox4) * C;
#pragma unroll
    for (int p = 0; p < 4; p++)
#pragma unroll
        for (int ch = 0; ch < C; ch++) o[p * C + ch] = (uint8_t)(((s[p][ch] + n / 2) * mult) >> 24);
    return true;
}

__global__ void __launch_bounds__(128) k_reduce(const PageD* __restrict__ pages) {
    const PageD& P = pages[blockIdx.z];
    if (!P.red) return;
    const int oy = blockIdx.y;
    if (oy >= P.rh) return;
    const int ox4 = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (ox4 >= P.rw) return;
    bool done = false;
    if (ox4 + 4 <= P.rw) {
        const int key = P.fx * 8 + P.pc;
        switch (key) {
            case 2 * 8 + 3: done = reduce_quad_words<2, 3>(P, ox4, oy); break;
            case 2 * 8 + 1: done = reduce_quad_words<2, 1>(P, ox4, oy); break;
            case 3 * 8 + 3: done = reduce_quad_words<3, 3>(P, ox4, oy); break;
            case 3 * 8 + 1: done = reduce_quad_words<3, 1>(P, ox4, oy); break;
            case 4 * 8 + 3: done = reduce_quad_words<4, 3>(P, ox4, oy); break;
            case 4 * 8 + 1: done = reduce_quad_words<4, 1>(P, ox4, oy); break;
            default: break;
        }
    }
    if (!done)
        for (int ox = ox4; ox < min(P.rw, ox4 + 4); ox++) reduce_pixel_bytes(P, ox, oy);
}

int launch_reduce(const PageD* d_pages, int npages, int max_rh, int max_rw, cudaStream_t st) {
    if (npages == 0 || max_rh == 0) return 0;
    dim3 grid(((max_rw + 3) / 4 + 127) / 128, max_rh, npages);
    k_reduce<<<grid, 128, 0, st>>>(d_pages);
    return 1;
}

// ------------------------------------------------------------------------------------------ resample
__device__ __forceinline__ uint8_t clip8(int v) {
    v >>= 22;
    return (uint8_t)min(max(v, 0), 255);
}

// Horizontal pass.  A CTA owns 128 output pixels x 8 input rows.  The input span those outputs read (bounds are
// monotonic: [xmin of the first, xmin+n of the last)) is staged row by row into shared memory with coalesced word
// loads; a thread keeps 8 rows x C accumulators in registers and walks the taps once, so every Q22 coefficient is
// loaded once per 8 rows and every pixel byte comes from shared memory.  Spans too long for the tile (extreme
// down-scales) take the direct path below.
constexpr int kHPix = 128;           // output pixels per CTA
constexpr int kHRows = 8;            // input rows per CTA
constexpr int kHSpanMax = 4096;      // staged bytes per row (aligned span)

template <int C>
__device__ __forceinline__ void resample_h_direct(const PageD& P, int y, int xx) {
    const int w = P.w;
    const int xmin = __ldg(P.hb + 2 * xx), n = __ldg(P.hb + 2 * xx + 1);
    const uint8_t* __restrict__ row = P.hin + (int64_t)y * P.hin_stride + (int64_t)xmin * C;
    uint8_t* __restrict__ out = P.tmp + (int64_t)y * P.tmp_stride + (int64_t)xx * C;
    int a[C];
#pragma unroll
    for (int ch = 0; ch < C; ch++) a[ch] = 1 << 21;
    for (int k = 0; k < n; k++) {
        const int kv = __ldg(P.hk + (int64_t)k * w + xx);
#pragma unroll
        for (int ch = 0; ch < C; ch++) a[ch] += (int)__ldg(row + C * k + ch) * kv;
    }
#pragma unroll
    for (int ch = 0; ch < C; ch++) out[ch] = clip8(a[ch]);
}

template <int C>
__device__ __forceinline__ void resample_h_tile(const PageD& P, uint32_t* sm) {
    const int w = P.w;
    const int xx0 = blockIdx.x * kHPix;
    const int y0 = blockIdx.y * kHRows;
    const int xx = xx0 + threadIdx.x;
    const int xxl = min(w, xx0 + kHPix) - 1;                       // last output pixel of the tile
    const int lo_px = __ldg(P.hb + 2 * xx0);
    const int hi_px = __ldg(P.hb + 2 * xxl) + __ldg(P.hb + 2 * xxl + 1);
    const int rows = min(kHRows, P.rh - y0);
    // byte span of the tile inside an input row, widened to aligned words of the row's global address
    const int64_t lo_b = (int64_t)lo_px * C, hi_b = (int64_t)hi_px * C;
    const uint8_t* row0 = P.hin + (int64_t)y0 * P.hin_stride;
    const bool rows_aligned = (P.hin_stride & 3) == 0;             // same (address & 3) for every row of the tile
    const int mis = (int)(((uintptr_t)row0 + lo_b) & 3);
    const int span_words = (int)((mis + (hi_b - lo_b) + 3) >> 2);
    if (!rows_aligned || span_words * 4 > kHSpanMax) {              // direct path (block-uniform decision)
        if (xx >= w) return;
        if (rows == kHRows) {
            // all 8 rows at once: a coefficient is loaded once per 8 rows, the pixel bytes come through L1 (rows of unaligned
            // stride, e.g. 2550 RGB pixels = 7650 bytes, cannot share one staged alignment)
            const int xmin = __ldg(P.hb + 2 * xx), n = __ldg(P.hb + 2 * xx + 1);
            const uint8_t* __restrict__ p = row0 + (int64_t)xmin * C;
            int acc[kHRows][C];
#pragma unroll
            for (int r = 0; r < kHRows; r++)
#pragma unroll
                for (int ch = 0; ch < C; ch++) acc[r][ch] = 1 << 21;
            for (int k = 0; k < n; k++) {
                const int kv = __ldg(P.hk + (int64_t)k * w + xx);
#pragma unroll
                for (int r = 0; r < kHRows; r++)
#pragma unroll
                    for (int ch = 0; ch < C; ch++) acc[r][ch] += (int)__ldg(p + (int64_t)r * P.hin_stride + C * k + ch) * kv;
            }
#pragma unroll
            for (int r = 0; r < kHRows; r++) {
                uint8_t* out = P.tmp + (int64_t)(y0 + r) * P.tmp_stride + (int64_t)xx * C;
#pragma unroll
                for (int ch = 0; ch < C; ch++) out[ch] = clip8(acc[r][ch]);
            }
            return;
        }
        for (int r = 0; r < rows; r++) resample_h_direct<C>(P, y0 + r, xx);
        return;
    }
    for (int r = 0; r < rows; r++) {
        const uint32_t* g = reinterpret_cast<const uint32_t*>(row0 + (int64_t)r * P.hin_stride + lo_b - mis);
        uint32_t* d = sm + r * (kHSpanMax / 4);
        for (int j = threadIdx.x; j < span_words; j += kHPix) d[j] = __ldg(g + j);
    }
    __syncthreads();
    if (xx >= w) return;
    const int xmin = __ldg(P.hb + 2 * xx), n = __ldg(P.hb + 2 * xx + 1);
    const uint8_t* sb = reinterpret_cast<const uint8_t*>(sm) + mis + (xmin - lo_px) * C;
    int acc[kHRows][C];
#pragma unroll
    for (int r = 0; r < kHRows; r++)
#pragma unroll
        for (int ch = 0; ch < C; ch++) acc[r][ch] = 1 << 21;
    for (int k = 0; k < n; k++) {
        const int kv = __ldg(P.hk + (int64_t)k * w + xx);
        const uint8_t* pk = sb + C * k;
#pragma unroll
        for (int r = 0; r < kHRows; r++)
#pragma unroll
            for (int ch = 0; ch < C; ch++) acc[r][ch] += (int)pk[r * kHSpanMax + ch] * kv;     // rows past `rows` read stale smem, never stored
    }
    for (int r = 0; r < rows; r++) {
        uint8_t* out = P.tmp + (int64_t)(y0 + r) * P.tmp_stride + (int64_t)xx * C;
#pragma unroll
        for (int ch = 0; ch < C; ch++) out[ch] = clip8(acc[r][ch]);
    }
}

// The fast path (every page-sized resize: up to 16 taps, 1 or 3 channels, span <= 4 KB).  The 8 input rows of the tile are fetched
// by TMA bulk copies — one elected thread arms an mbarrier with the byte count and issues one copy per row from the row's address
// rounded down to 16 bytes, whatever the row stride — so a row lands in shared memory with its own byte offset.  A thread owns one
// output pixel: it keeps its 16 Q22 coefficients in registers, and per row reads the 13 (5) words that hold its taps, realigns
// them with funnel shifts, and multiplies byte lanes out of registers (PRMT + IMAD): 13 shared loads per row instead of 45 byte
// loads through L1 (the load/store unit, not the multiplier, was the bound of the byte version).
constexpr int kHTaps = 16;
constexpr int kHRowBuf = kHSpanMax + 64;       // staged row: span rounded to 16 at both ends + the words a thread reads past its taps

template <int C>
__device__ __forceinline__ void resample_h_tma(const PageD& P, uint8_t (*tile)[kHRowBuf], uint64_t* bar,
                                               int lo_px, int span_bytes, int y0, int rows) {
    const int w = P.w;
    const int xx = blockIdx.x * kHPix + threadIdx.x;
    const uint8_t* row0 = P.hin + (int64_t)y0 * P.hin_stride + (int64_t)lo_px * C;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        uint32_t total = 0;
        for (int r = 0; r < rows; r++) {
            const uintptr_t a = (uintptr_t)(row0 + (int64_t)r * P.hin_stride);
            total += (uint32_t)((((a & 15) + span_bytes) + 15) & ~15);
        }
        mbar_expect_tx(bar, total);
        for (int r = 0; r < rows; r++) {
            const uintptr_t a = (uintptr_t)(row0 + (int64_t)r * P.hin_stride);
            bulk_g2s(tile[r], reinterpret_cast<const void*>(a & ~(uintptr_t)15), (uint32_t)((((a & 15) + span_bytes) + 15) & ~15), bar);
        }
    }
    // coefficients while the rows are in flight
    int kv[kHTaps];
    int xmin = lo_px;
    if (xx < w) {
        xmin = __ldg(P.hb + 2 * xx);
#pragma unroll
        for (int k = 0; k < kHTaps; k++) kv[k] = k < P.hks ? __ldg(P.hk + (int64_t)k * w + xx) : 0;   // rows of the table are zero padded past n
    }
    __syncthreads();                           // the barrier is initialised before anyone waits on it
    mbar_wait(bar, 0);
    if (xx >= w) return;
    constexpr int NW = (kHTaps * C + 3) / 4;   // words that hold the taps of one row
    for (int r = 0; r < rows; r++) {
        const int o = (int)((uintptr_t)(row0 + (int64_t)r * P.hin_stride) & 15) + (xmin - lo_px) * C;
        const uint32_t* tw = reinterpret_cast<const uint32_t*>(tile[r]) + (o >> 2);
        const int sh = (o & 3) * 8;
        uint32_t wd[NW + 1];
#pragma unroll
        for (int i = 0; i <= NW; i++) wd[i] = tw[i];
#pragma unroll
        for (int i = 0; i < NW; i++) wd[i] = __funnelshift_r(wd[i], wd[i + 1], sh);
        int acc[C];
#pragma unroll
        for (int ch = 0; ch < C; ch++) acc[ch] = 1 << 21;
#pragma unroll
        for (int k = 0; k < kHTaps; k++)
#pragma unroll
            for (int ch = 0; ch < C; ch++) {
                const int b = k * C + ch;
                acc[ch] += (int)__byte_perm(wd[b >> 2], 0u, 0x4440u | (uint32_t)(b & 3)) * kv[k];
            }
        uint8_t* out = P.tmp + (int64_t)(y0 + r) * P.tmp_stride + (int64_t)xx * C;
#pragma unroll
        for (int ch = 0; ch < C; ch++) out[ch] = clip8(acc[ch]);
    }
}

__global__ void __launch_bounds__(kHPix) k_resample_h(const PageD* __restrict__ pages) {
    __shared__ __align__(128) uint8_t tile[kHRows][kHRowBuf];
    __shared__ __align__(8) uint64_t bar;
    const PageD& P = pages[blockIdx.z];
    if (!P.tmp) return;
    if ((int)blockIdx.y * kHRows >= P.rh || (int)blockIdx.x * kHPix >= P.w) return;
    {
        const int xx0 = blockIdx.x * kHPix, xxl = min(P.w, xx0 + kHPix) - 1, y0 = blockIdx.y * kHRows;
        const int lo_px = __ldg(P.hb + 2 * xx0), hi_px = __ldg(P.hb + 2 * xxl) + __ldg(P.hb + 2 * xxl + 1);
        const int span = (hi_px - lo_px) * P.pc;
        if (P.hks <= kHTaps && span + 32 <= kHSpanMax) {               // block-uniform
            const int rows = min(kHRows, P.rh - y0);
            if (P.pc == 3) resample_h_tma<3>(P, tile, &bar, lo_px, span, y0, rows);
            else resample_h_tma<1>(P, tile, &bar, lo_px, span, y0, rows);
            return;
        }
    }
    uint32_t* sm = reinterpret_cast<uint32_t*>(&tile[0][0]);          // long kernels (extreme down-scales): the word-staged / direct paths
    if (P.pc == 3) resample_h_tile<3>(P, sm); else resample_h_tile<1>(P, sm);
}

// Vertical pass.  The tap weight is the same for a whole output row, the taps are whole input rows: a thread owns 16 consecutive
// bytes of one output row — one 128-bit load per tap, sixteen int32 accumulators — when rows are 16-byte aligned (always inside the
// arena: the horizontal pass writes rows padded to 16), 4 bytes when they are only word aligned, otherwise one byte.
__device__ __forceinline__ uint32_t pack_clip4(const int* a) {
    return (uint32_t)clip8(a[0]) | ((uint32_t)clip8(a[1]) << 8) | ((uint32_t)clip8(a[2]) << 16) | ((uint32_t)clip8(a[3]) << 24);
}

__global__ void __launch_bounds__(128) k_resample_v(const PageD* __restrict__ pages) {
    const PageD& P = pages[blockIdx.z];
    if (!P.vout) return;
    const int yy = blockIdx.y;
    if (yy >= P.h) return;
    const int wc = P.w * P.pc;
    const int ymin = __ldg(P.vb + 2 * yy), n = __ldg(P.vb + 2 * yy + 1);
    const uintptr_t geo = ((uintptr_t)P.vin) | (uintptr_t)P.vin_stride | ((uintptr_t)P.vout) | (uintptr_t)P.vout_stride;
    const int i16 = blockIdx.x * blockDim.x + threadIdx.x;      // 16-byte group in the row
    if (16 * i16 >= wc) return;
    const int32_t* __restrict__ kv = P.vk + yy;                  // coefficient of tap k at kv[k * h]
    if ((geo & 15) == 0) {
        // padded rows: the bytes past wc inside the last group are row padding, reading and writing them is harmless
        const uint4* __restrict__ col = reinterpret_cast<const uint4*>(P.vin + (int64_t)ymin * P.vin_stride) + i16;
        const int64_t sw = P.vin_stride >> 4;
        int a[16];
#pragma unroll
        for (int j = 0; j < 16; j++) a[j] = 1 << 21;
        for (int k = 0; k < n; k++) {
            const uint4 v = __ldg(col + (int64_t)k * sw);
            const int c = __ldg(kv + (int64_t)k * P.h);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 16; j++) a[j] += (int)__byte_perm(w[j >> 2], 0u, 0x4440u | (uint32_t)(j & 3)) * c;
        }
        *(reinterpret_cast<uint4*>(P.vout + (int64_t)yy * P.vout_stride) + i16) =
            make_uint4(pack_clip4(a), pack_clip4(a + 4), pack_clip4(a + 8), pack_clip4(a + 12));
    } else if ((geo & 3) == 0) {
        for (int i4 = 4 * i16; i4 < 4 * i16 + 4 && 4 * i4 < wc; i4++) {
            const uint32_t* __restrict__ col = reinterpret_cast<const uint32_t*>(P.vin + (int64_t)ymin * P.vin_stride) + i4;
            const int64_t sw = P.vin_stride >> 2;
            int a[4] = {1 << 21, 1 << 21, 1 << 21, 1 << 21};
            for (int k = 0; k < n; k++) {
                const uint32_t v = __ldg(col + (int64_t)k * sw);
                const int c = __ldg(kv + (int64_t)k * P.h);
#pragma unroll
                for (int j = 0; j < 4; j++) a[j] += (int)__byte_perm(v, 0u, 0x4440u | (uint32_t)j) * c;
            }
            reinterpret_cast<uint32_t*>(P.vout + (int64_t)yy * P.vout_stride)[i4] = pack_clip4(a);   // row padding absorbs the tail
        }
    } else {
        for (int i = 16 * i16; i < min(wc, 16 * i16 + 16); i++) {
            const uint8_t* __restrict__ col = P.vin + (int64_t)ymin * P.vin_stride + i;
            int a = 1 << 21;
            for (int k = 0; k < n; k++) a += (int)__ldg(col + (int64_t)k * P.vin_stride) * __ldg(kv + (int64_t)k * P.h);
            P.vout[(int64_t)yy * P.vout_stride + i] = clip8(a);
        }
    }
}

int launch_resample_h(const PageD* d_pages, int npages, int max_rh, int max_w, cudaStream_t st) {
    if (npages == 0 || max_rh == 0) return 0;
    dim3 grid((max_w + kHPix - 1) / kHPix, (max_rh + kHRows - 1) / kHRows, npages);
    k_resample_h<<<grid, kHPix, 0, st>>>(d_pages);
    return 1;
}

int launch_resample_v(const PageD* d_pages, int npages, int max_h, int max_wc, cudaStream_t st) {
    if (npages == 0 || max_h == 0) return 0;
    dim3 grid(((max_wc + 15) / 16 + 127) / 128, max_h, npages);
    k_resample_v<<<grid, 128, 0, st>>>(d_pages);
    return 1;
}

// ------------------------------------------------------------------------------------------ coefficients (host)
static inline double f_sinc(double x) { if (x == 0.0) return 1.0; x *= M_PI; return sin(x) / x; }
static double flt_box(double x) { return (x > -0.5 && x <= 0.5) ? 1.0 : 0.0; }
static double flt_bilinear(double x) { if (x < 0.0) x = -x; return x < 1.0 ? 1.0 - x : 0.0; }
static double flt_hamming(double x) {
    if (x < 0.0) x = -x;
    if (x == 0.0) return 1.0;
    if (x >= 1.0) return 0.0;
    x *= M_PI;
    return sin(x) / x * (0.54 + 0.46 * cos(x));
}
static double flt_bicubic(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}
static double flt_lanczos(double x) { return (x >= -3.0 && x < 3.0) ? f_sinc(x) * f_sinc(x / 3) : 0.0; }

// Window [xmin, xmin+n) and n Q22 weights per output index; kk is out_size x ksize (row-major, zero padded).
int resample_coeffs_host(int in_size, int out_size, int filter, float box0, float box1,
                         int32_t* bounds, int32_t* kk, int* ksize_out) {
    double (*fn)(double); double fsupport;
    switch (filter) {
        case 1: fn = flt_lanczos; fsupport = 3.0; break;
        case 2: fn = flt_bilinear; fsupport = 1.0; break;
        case 3: fn = flt_bicubic; fsupport = 2.0; break;
        case 4: fn = flt_box; fsupport = 0.5; break;
        case 5: fn = flt_hamming; fsupport = 1.0; break;
        default: return -1;
    }
    if (in_size <= 0 || out_size <= 0) return -1;
    const double in0 = (double)box0, in1 = (double)box1;
    const double scale = (in1 - in0) / out_size;
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = fsupport * filterscale;
    const int ksize = (int)ceil(support) * 2 + 1;
    if (ksize_out) *ksize_out = ksize;
    if (!kk || !bounds) return 0;
    const double ss = 1.0 / filterscale;
    std::vector<double> w(ksize);
    for (int xx = 0; xx < out_size; xx++) {
        const double center = in0 + (xx + 0.5) * scale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        const int n = xmax - xmin;
        double ww = 0.0;
        for (int x = 0; x < n; x++) { w[x] = fn((x + xmin - center + 0.5) * ss); ww += w[x]; }
        int32_t* k = kk + (int64_t)xx * ksize;
        for (int x = 0; x < ksize; x++) k[x] = 0;
        for (int x = 0; x < n; x++) {
            double v = w[x];
            if (ww != 0.0) v /= ww;
            k[x] = v < 0 ? (int32_t)(-0.5 + v * (1 << 22)) : (int32_t)(0.5 + v * (1 << 22));
        }
        bounds[2 * xx] = xmin; bounds[2 * xx + 1] = n;
    }
    return 0;
}

}  // namespace vcp
