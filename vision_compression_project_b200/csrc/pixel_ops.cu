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
__global__ void k_convert(const PageD* __restrict__ pages) {
    const PageD& P = pages[blockIdx.z];
    if (!P.conv) return;
    const int y = blockIdx.y;
    if (y >= P.sh) return;
    const int sc = P.sc, c = P.pc;
    const uint8_t* __restrict__ srow = P.src + (int64_t)y * P.src_stride;
    uint8_t* __restrict__ drow = P.conv + (int64_t)y * P.sw * c;
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < P.sw; x += gridDim.x * blockDim.x) {
        const uint8_t* s = srow + (int64_t)x * sc;
        uint8_t* d = drow + (int64_t)x * c;
        if (c == 3) {
            if (sc <= 2) { uint8_t v = __ldg(s); d[0] = v; d[1] = v; d[2] = v; }            // L, LA -> RGB (alpha dropped)
            else { d[0] = __ldg(s); d[1] = __ldg(s + 1); d[2] = __ldg(s + 2); }             // RGB(A) -> RGB
        } else {                                                                            // -> L
            if (sc <= 2) d[0] = __ldg(s);
            else {
                uint32_t r = __ldg(s), g = __ldg(s + 1), b = __ldg(s + 2);
                d[0] = (uint8_t)((r * 19595u + g * 38470u + b * 7471u + 0x8000u) >> 16);
            }
        }
    }
}

int launch_convert(const PageD* d_pages, int npages, int max_rows, int max_w, cudaStream_t st) {
    if (npages == 0 || max_rows == 0) return 0;
    dim3 grid((max_w + 255) / 256, max_rows, npages);
    k_convert<<<grid, 256, 0, st>>>(d_pages);
    return 1;
}

// ------------------------------------------------------------------------------------------ reduce
__global__ void k_reduce(const PageD* __restrict__ pages) {
    const PageD& P = pages[blockIdx.z];
    if (!P.red) return;
    const int oy = blockIdx.y;
    if (oy >= P.rh) return;
    const int c = P.pc, fx = P.fx, fy = P.fy;
    const int y0 = oy * fy, y1 = min(y0 + fy, P.sh);
    for (int ox = blockIdx.x * blockDim.x + threadIdx.x; ox < P.rw; ox += gridDim.x * blockDim.x) {
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
}

int launch_reduce(const PageD* d_pages, int npages, int max_rh, int max_rw, cudaStream_t st) {
    if (npages == 0 || max_rh == 0) return 0;
    dim3 grid((max_rw + 127) / 128, max_rh, npages);
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

__global__ void __launch_bounds__(kHPix) k_resample_h(const PageD* __restrict__ pages) {
    __shared__ __align__(16) uint32_t sm[kHRows * kHSpanMax / 4];
    const PageD& P = pages[blockIdx.z];
    if (!P.tmp) return;
    if ((int)blockIdx.y * kHRows >= P.rh || (int)blockIdx.x * kHPix >= P.w) return;
    if (P.pc == 3) resample_h_tile<3>(P, sm); else resample_h_tile<1>(P, sm);
}

// Vertical pass: a thread owns 4 consecutive bytes of one output row (one aligned word per tap) when the rows are
// word aligned (always inside the arena), otherwise one byte.  The tap weight is the same for the whole row.
__global__ void __launch_bounds__(256) k_resample_v(const PageD* __restrict__ pages) {
    const PageD& P = pages[blockIdx.z];
    if (!P.vout) return;
    const int yy = blockIdx.y;
    if (yy >= P.h) return;
    const int wc = P.w * P.pc;
    const int ymin = __ldg(P.vb + 2 * yy), n = __ldg(P.vb + 2 * yy + 1);
    const bool aligned = ((((uintptr_t)P.vin) | (uintptr_t)P.vin_stride | ((uintptr_t)P.vout) | (uintptr_t)P.vout_stride) & 3) == 0;
    const int i4 = blockIdx.x * blockDim.x + threadIdx.x;       // word index in the row
    if (aligned) {
        if (4 * i4 >= wc) return;
        const uint32_t* __restrict__ col = reinterpret_cast<const uint32_t*>(P.vin + (int64_t)ymin * P.vin_stride) + i4;
        const int64_t sw = P.vin_stride >> 2;
        int a0 = 1 << 21, a1 = 1 << 21, a2 = 1 << 21, a3 = 1 << 21;
        for (int k = 0; k < n; k++) {
            const uint32_t v = __ldg(col + (int64_t)k * sw);
            const int kv = __ldg(P.vk + (int64_t)k * P.h + yy);
            a0 += (int)(v & 0xFFu) * kv; a1 += (int)((v >> 8) & 0xFFu) * kv; a2 += (int)((v >> 16) & 0xFFu) * kv; a3 += (int)(v >> 24) * kv;
        }
        const uint32_t o = (uint32_t)clip8(a0) | ((uint32_t)clip8(a1) << 8) | ((uint32_t)clip8(a2) << 16) | ((uint32_t)clip8(a3) << 24);
        // padded rows: the bytes past wc inside the last word are row padding, writing them is harmless
        reinterpret_cast<uint32_t*>(P.vout + (int64_t)yy * P.vout_stride)[i4] = o;
    } else {
        for (int i = 4 * i4; i < min(wc, 4 * i4 + 4); i++) {
            const uint8_t* __restrict__ col = P.vin + (int64_t)ymin * P.vin_stride + i;
            int a = 1 << 21;
            for (int k = 0; k < n; k++) a += (int)__ldg(col + (int64_t)k * P.vin_stride) * __ldg(P.vk + (int64_t)k * P.h + yy);
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
    dim3 grid(((max_wc + 3) / 4 + 255) / 256, max_h, npages);
    k_resample_v<<<grid, 256, 0, st>>>(d_pages);
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
