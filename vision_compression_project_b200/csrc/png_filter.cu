// png_filter.cu — PNG adaptive row filtering, bit-exact with Pillow's ZipEncode.c rule, plus the
// Adler-32 partials of the filtered stream.
//
// Replaces the per-row part of `page_image.save(path)` (backend/app/pipeline/pdf_extract.py:130):
// Pack.c (RGBX -> packed row) + ZipEncode.c (candidates None/Up/Sub/[Avg]/Paeth, pick the smallest
// sum of |signed residual| in the order Up, Sub, Avg(optimize only), Paeth with a strict '<' and an
// early-out once a sum is 0).  Restated in oracle/restate.py:filter_row and pinned by the 23 recorded
// PNGs of the reference (tests/golden/fixtures.json: 2339 filter decisions per page).
//
// One CTA per image row.  Both rows are staged once in shared memory with 4-byte loads on the
// aligned-down address, every candidate is evaluated on 4 bytes per instruction with the byte-SIMD
// intrinsics (no per-byte loop), the five sums are reduced with warp shuffles, and only the winning
// residual is materialised.  The output row (1 + W*bpp bytes, almost never 4-byte aligned) is staged
// in shared memory at the same alignment as its global address so the body leaves as aligned u32 stores.
// HBM-bound by design: algorithmic traffic = read W*bpp + write 1+W*bpp per row (the previous row is
// re-read through L2).
#include "vcp_internal.cuh"

namespace vcp {

namespace {

constexpr int kThreads = 128;

__device__ __forceinline__ uint32_t ld4(const uint32_t* sm, int byte_off) {   // unaligned 4-byte read from shared
    const int w = byte_off >> 2;
    return __funnelshift_r(sm[w], sm[w + 1], (byte_off & 3) * 8);
}

__device__ __forceinline__ uint32_t paeth4(uint32_t a, uint32_t b, uint32_t c) {
    const uint32_t pa = __vabsdiffu4(b, c), pb = __vabsdiffu4(a, c);
    const uint32_t same = ~(__vcmpgeu4(a, c) ^ __vcmpgeu4(b, c));          // (a-c) and (b-c) have the same sign
    const uint32_t pc = (same & __vaddus4(pa, pb)) | (~same & __vabsdiffu4(pa, pb));
    const uint32_t sa = __vcmpleu4(pa, pb) & __vcmpleu4(pa, pc);
    const uint32_t sb = ~sa & __vcmpleu4(pb, pc);
    return (a & sa) | (b & sb) | (c & ~(sa | sb));
}

__device__ __forceinline__ uint32_t score4(uint32_t v) { return __vsadu4(__vabs4(v), 0u); }

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Stage `nbytes` starting at global `g` (any alignment) into shared words sm[1..], so that row byte i sits at
// shared byte offset 4 + o + i (o = g & 3).  sm[0] and the o bytes in front of the row are zero ("left of the
// first pixel"), words past the row are zero-filled up to `total_words`.
__device__ __forceinline__ void stage_row(uint32_t* sm, const uint8_t* g, int nbytes, int total_words, bool zero_row) {
    const int o = (int)((uintptr_t)g & 3);
    const uint32_t* ga = (const uint32_t*)(g - o);
    const int m = (o + nbytes + 3) >> 2;                    // aligned words that contain row bytes
    for (int j = threadIdx.x; j < total_words; j += kThreads) {
        uint32_t v = 0;
        if (!zero_row && j >= 1 && j <= m) {
            v = __ldg(ga + (j - 1));
            if (j == 1 && o) v &= 0xFFFFFFFFu << (8 * o);  // bytes in front of the row
            if (j == m) { const int keep = (o + nbytes) - 4 * (m - 1); if (keep < 4) v &= 0xFFFFFFFFu >> (8 * (4 - keep)); }
        }
        sm[j] = v;
    }
}

}  // namespace

// dynamic smem: cur[words] | prv[words] | out[words], words = (max_rowbytes + 16) / 4 + 4
__global__ void __launch_bounds__(kThreads) k_png_filter(const PageD* __restrict__ pages, int optimize,
                                                         uint32_t* __restrict__ row_adler, int words) {
    extern __shared__ uint32_t smem[];
    __shared__ uint32_t red[5][kThreads / 32];
    __shared__ uint32_t red2[2][kThreads / 32];
    const PageD& P = pages[blockIdx.y];
    const int y = blockIdx.x;
    if (y >= P.h) return;
    const int bpp = P.c;
    const int n = P.w * bpp;                                  // row bytes
    uint32_t* smc = smem; uint32_t* smp = smem + words; uint32_t* smo = smem + 2 * words;
    const uint8_t* grow = P.pix + (int64_t)y * P.pix_stride;
    const uint8_t* gprev = grow - P.pix_stride;
    const int oc = (int)((uintptr_t)grow & 3), op = (int)((uintptr_t)gprev & 3);
    stage_row(smc, grow, n, words, false);
    stage_row(smp, gprev, n, words, y == 0);
    __syncthreads();
    const int bc = 4 + oc, bp = 4 + op;                       // shared byte offset of row byte 0
    const int ngroups = (n + 3) >> 2;

    // ---- pass 1: the five candidate sums
    uint32_t s_none = 0, s_up = 0, s_sub = 0, s_avg = 0, s_pae = 0;
    for (int g = threadIdx.x; g < ngroups; g += kThreads) {
        const int i = 4 * g;
        const uint32_t cur = ld4(smc, bc + i), up = ld4(smp, bp + i);
        const uint32_t left = ld4(smc, bc + i - bpp), ul = ld4(smp, bp + i - bpp);
        uint32_t mask = 0xFFFFFFFFu;
        if (n - i < 4) mask >>= 8 * (4 - (n - i));
        s_none += score4(cur & mask);
        s_up += score4(__vsub4(cur, up) & mask);
        s_sub += score4(__vsub4(cur, left) & mask);
        if (optimize) s_avg += score4(__vsub4(cur, __vhaddu4(left, up)) & mask);
        s_pae += score4(__vsub4(cur, paeth4(left, up, ul)) & mask);
    }
    s_none = warp_sum(s_none); s_up = warp_sum(s_up); s_sub = warp_sum(s_sub); s_avg = warp_sum(s_avg); s_pae = warp_sum(s_pae);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[0][warp] = s_none; red[1][warp] = s_up; red[2][warp] = s_sub; red[3][warp] = s_avg; red[4][warp] = s_pae; }
    __syncthreads();
    uint32_t tot[5];
#pragma unroll
    for (int k = 0; k < 5; k++) { uint32_t t = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; w++) t += red[k][w]; tot[k] = t; }
    // ZipEncode.c order: None is the incumbent; Up, Sub, (Avg), Paeth replace it only if strictly smaller,
    // and nothing is tried once the incumbent's sum is 0.
    int ftype = 0; uint32_t best = tot[0];
    if (best > 0 && tot[1] < best) { best = tot[1]; ftype = 2; }
    if (best > 0 && tot[2] < best) { best = tot[2]; ftype = 1; }
    if (optimize && best > 0 && tot[3] < best) { best = tot[3]; ftype = 3; }
    if (best > 0 && tot[4] < best) { best = tot[4]; ftype = 4; }

    // ---- pass 2: materialise the winner into the output staging row, Adler partials on the way
    const int64_t L = (int64_t)n + 1;                         // output row bytes
    uint8_t* gout = P.filt + (int64_t)y * L;
    const int a0 = (int)((uintptr_t)gout & 3);
    uint8_t* so = (uint8_t*)smo;
    uint32_t a_s1 = 0; uint64_t a_s2 = 0;                     // sum b, sum (L-k)*b over this thread's bytes (k = index in out row)
    if (threadIdx.x == 0) { so[a0] = (uint8_t)ftype; a_s1 = ftype; a_s2 = (uint64_t)L * ftype; }
    for (int g = threadIdx.x; g < ngroups; g += kThreads) {
        const int i = 4 * g;
        const uint32_t cur = ld4(smc, bc + i);
        uint32_t r;
        if (ftype == 0) r = cur;
        else if (ftype == 2) r = __vsub4(cur, ld4(smp, bp + i));
        else if (ftype == 1) r = __vsub4(cur, ld4(smc, bc + i - bpp));
        else if (ftype == 3) r = __vsub4(cur, __vhaddu4(ld4(smc, bc + i - bpp), ld4(smp, bp + i)));
        else r = __vsub4(cur, paeth4(ld4(smc, bc + i - bpp), ld4(smp, bp + i), ld4(smp, bp + i - bpp)));
        const int valid = min(4, n - i);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (k < valid) {
                const uint32_t b = (r >> (8 * k)) & 0xFFu;
                so[a0 + 1 + i + k] = (uint8_t)b;
                a_s1 += b; a_s2 += (uint64_t)(L - 1 - i - k) * b;
            }
        }
    }
    a_s1 %= 65521u;
    uint32_t a_s2m = (uint32_t)(a_s2 % 65521u);
    a_s1 = warp_sum(a_s1); a_s2m = warp_sum(a_s2m);
    if (lane == 0) { red2[0][warp] = a_s1; red2[1][warp] = a_s2m; }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t1 = 0, t2 = 0;
        for (int w = 0; w < kThreads / 32; w++) { t1 += red2[0][w]; t2 += red2[1][w]; }
        row_adler[P.row0 + y] = ((t2 % 65521u) << 16) | (t1 % 65521u);
    }
    // ---- copy out: aligned words of [a0, a0+L) as u32, ragged head/tail as bytes
    const int first_w = (a0 + 3) >> 2;                        // first fully covered word
    const int end_w = (int)((a0 + L) >> 2);                   // one past the last fully covered word
    uint32_t* gw = (uint32_t*)(gout - a0);
    for (int w = first_w + threadIdx.x; w < end_w; w += kThreads) gw[w] = smo[w];
    if (threadIdx.x < 4) {
        const int k = threadIdx.x;                            // head bytes: output offsets [0, 4*first_w - a0)
        if (a0 && a0 + k < 4 && k < L && first_w * 4 - a0 > k) gout[k] = so[a0 + k];
    } else if (threadIdx.x < 8) {
        const int k = threadIdx.x - 4;                        // tail bytes: shared offsets [4*end_w, a0+L)
        const int64_t sb = (int64_t)end_w * 4 + k;
        if (end_w >= first_w && sb < a0 + L) gout[sb - a0] = so[sb];
    }
    if (end_w < first_w) {                                    // row shorter than one aligned word: plain bytes
        for (int k = threadIdx.x; k < L; k += kThreads) gout[k] = so[a0 + k];
    }
}

int launch_png_filter(const PageD* d_pages, int npages, int max_h, int max_rowbytes, int optimize,
                         uint32_t* row_adler, cudaStream_t st) {
    if (npages == 0 || max_h == 0) return 0;
    const int words = (max_rowbytes + 16) / 4 + 4;
    const size_t smem = (size_t)words * 3 * sizeof(uint32_t);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        cudaFuncSetAttribute(k_png_filter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = smem;
    }
    dim3 grid(max_h, npages);
    k_png_filter<<<grid, kThreads, smem, st>>>(d_pages, optimize, row_adler, words);
    return 1;
}

// ---------------------------------------------------------------------------------------- Adler-32 combine
// One thread per page walks its rows: after a row of L bytes with partials (s1, s2):
//   b += L*a + s2 ; a += s1   (mod 65521)      [zlib adler32.c semantics, a=1,b=0 at the start]
__global__ void k_adler_combine(const PageD* __restrict__ pages, int npages, const uint32_t* __restrict__ row_adler,
                                uint32_t* __restrict__ page_adler) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npages) return;
    const PageD& P = pages[p];
    const uint32_t L = (uint32_t)(((int64_t)P.w * P.c + 1) % 65521);
    uint32_t a = 1, b = 0;
    for (int y = 0; y < P.h; y++) {
        const uint32_t v = __ldg(row_adler + P.row0 + y);
        b = (uint32_t)((b + (uint64_t)L * a + (v >> 16)) % 65521u);
        a = (a + (v & 0xFFFFu)) % 65521u;
    }
    page_adler[p] = (b << 16) | a;
}

int launch_adler_combine(const PageD* d_pages, int npages, const uint32_t* row_adler, uint32_t* page_adler, cudaStream_t st) {
    if (npages == 0) return 0;
    k_adler_combine<<<(npages + 63) / 64, 64, 0, st>>>(d_pages, npages, row_adler, page_adler);
    return 1;
}

}  // namespace vcp
