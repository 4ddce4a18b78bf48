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
    const int sc = P.sc, c = P.c;
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
    const int c = P.c, fx = P.fx, fy = P.fy;
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

// Horizontal pass: thread = one output pixel (all channels) of one row.
__global__ void k_resample_h(const PageD* __restrict__ pages) {
    const PageD& P = pages[blockIdx.z];
    if (!P.tmp) return;
    const int y = blockIdx.y;
    const int xx = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= P.rh || xx >= P.w) return;
    const int c = P.c, w = P.w;
    const int xmin = __ldg(P.hb + 2 * xx), n = __ldg(P.hb + 2 * xx + 1);
    const uint8_t* __restrict__ row = P.hin + (int64_t)y * P.hin_stride + (int64_t)xmin * c;
    uint8_t* __restrict__ out = P.tmp + ((int64_t)y * w + xx) * c;
    if (c == 3) {
        int a0 = 1 << 21, a1 = 1 << 21, a2 = 1 << 21;
        for (int k = 0; k < n; k++) {
            const int kv = __ldg(P.hk + (int64_t)k * w + xx);
            a0 += (int)__ldg(row + 3 * k) * kv; a1 += (int)__ldg(row + 3 * k + 1) * kv; a2 += (int)__ldg(row + 3 * k + 2) * kv;
        }
        out[0] = clip8(a0); out[1] = clip8(a1); out[2] = clip8(a2);
    } else {
        for (int ch = 0; ch < c; ch++) {
            int a = 1 << 21;
            for (int k = 0; k < n; k++) a += (int)__ldg(row + (int64_t)k * c + ch) * __ldg(P.hk + (int64_t)k * w + xx);
            out[ch] = clip8(a);
        }
    }
}

// Vertical pass: thread = one output byte (x*c + ch) of one output row; reads are contiguous along the row.
__global__ void k_resample_v(const PageD* __restrict__ pages) {
    const PageD& P = pages[blockIdx.z];
    if (!P.vout) return;
    const int yy = blockIdx.y;
    const int wc = P.w * P.c;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (yy >= P.h || i >= wc) return;
    const int ymin = __ldg(P.vb + 2 * yy), n = __ldg(P.vb + 2 * yy + 1);
    const uint8_t* __restrict__ col = P.vin + (int64_t)ymin * P.vin_stride + i;
    int a = 1 << 21;
    for (int k = 0; k < n; k++) a += (int)__ldg(col + (int64_t)k * P.vin_stride) * __ldg(P.vk + (int64_t)k * P.h + yy);
    P.vout[(int64_t)yy * wc + i] = clip8(a);
}

int launch_resample_h(const PageD* d_pages, int npages, int max_rh, int max_w, cudaStream_t st) {
    if (npages == 0 || max_rh == 0) return 0;
    dim3 grid((max_w + 127) / 128, max_rh, npages);
    k_resample_h<<<grid, 128, 0, st>>>(d_pages);
    return 1;
}

int launch_resample_v(const PageD* d_pages, int npages, int max_h, int max_wc, cudaStream_t st) {
    if (npages == 0 || max_h == 0) return 0;
    dim3 grid((max_wc + 255) / 256, max_h, npages);
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
