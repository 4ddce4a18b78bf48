// png_filter.cu — PNG adaptive row filtering, bit-exact with Pillow's ZipEncode.c rule, plus the
// Adler-32 partials of the filtered stream.
//
// Replaces the per-row part of `page_image.save(path)` (backend/app/pipeline/pdf_extract.py:130):
// Pack.c (RGBX -> packed row) + ZipEncode.c (candidates None/Up/Sub/[Avg]/Paeth, pick the smallest
// sum of |signed residual| in the order Up, Sub, Avg(optimize only), Paeth with a strict '<' and an
// early-out once a sum is 0).  Restated in oracle/restate.py:filter_row and pinned by the 23 recorded
// PNGs of the reference (tests/golden/fixtures.json: 2339 filter decisions per page).
//
// One CTA walks 8 consecutive rows of a page, so every row is read from global memory once (the previous row
// stays in shared memory).  A thread owns 16 contiguous row bytes: one 128-bit + two 32-bit shared loads per
// row give it the current/left and up/up-left bytes, every candidate is evaluated 4 bytes per instruction
// (VABSDIFF4 + dp4a for the |signed| sums, byte-SIMD emulation for Paeth), the sums are reduced with warp
// shuffles, and only the winning residual is materialised — as whole words in shared memory, shifted into
// place (the filter byte makes the output row 1 + W*bpp long, so it is never aligned with its input) and
// stored as aligned 32-bit words.  A row identical to the one above (blank paper) skips the candidates.
// HBM-bound by design: algorithmic traffic = read W*bpp + write 1 + W*bpp per row.
#include "vcp_internal.cuh"
#include <atomic>

namespace vcp {

namespace {

constexpr int kThreads = 128;
constexpr int kRowsPerCta = 8;
constexpr int kLead = 16;        // row byte 0 sits at shared byte kLead + (global address & 3): 16-byte chunks start word-aligned

__device__ __forceinline__ uint32_t paeth4(uint32_t a, uint32_t b, uint32_t c) {
    const uint32_t pa = __vabsdiffu4(b, c), pb = __vabsdiffu4(a, c);
    const uint32_t same = ~(__vcmpgeu4(a, c) ^ __vcmpgeu4(b, c));          // (a-c) and (b-c) have the same sign
    const uint32_t pc = (same & __vaddus4(pa, pb)) | (~same & __vabsdiffu4(pa, pb));
    const uint32_t sa = __vcmpleu4(pa, pb) & __vcmpleu4(pa, pc);
    const uint32_t sb = ~sa & __vcmpleu4(pb, pc);
    return (a & sa) | (b & sb) | (c & ~(sa | sb));
}

// sum over the 4 bytes of |signed byte| (0x80 counts 128), accumulated
__device__ __forceinline__ uint32_t score_acc(uint32_t v, uint32_t acc) {
    return __dp4a(__vabsdiffu4(v ^ 0x80808080u, 0x80808080u), 0x01010101u, acc);
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Stage `nbytes` starting at global `g` (any alignment) into shared words so that row byte i sits at shared byte
// kLead + o + i (o = g & 3).  Everything outside the row inside [0, total_words) is zero ("left of the first pixel").
__device__ __forceinline__ void stage_row(uint32_t* sm, const uint8_t* g, int nbytes, int total_words, bool zero_row) {
    const int o = (int)((uintptr_t)g & 3);
    const uint32_t* ga = (const uint32_t*)(g - o);
    const int m = (o + nbytes + 3) >> 2;                    // aligned words that contain row bytes
    constexpr int L = kLead / 4;
    for (int j = threadIdx.x; j < total_words; j += kThreads) {
        uint32_t v = 0;
        if (!zero_row && j >= L && j < L + m) {
            v = __ldg(ga + (j - L));
            if (j == L && o) v &= 0xFFFFFFFFu << (8 * o);  // bytes in front of the row
            if (j == L + m - 1) { const int keep = (o + nbytes) - 4 * (m - 1); if (keep < 4) v &= 0xFFFFFFFFu >> (8 * (4 - keep)); }
        }
        sm[j] = v;
    }
}

// Gray -> RGB on the way in (Convert.c L/LA -> RGB is plain replication): `g` is a row of wpix gray bytes, the staged row is
// the 3*wpix interleaved bytes Pillow would have produced, at shared byte kLead (alignment offset 0).
__device__ __forceinline__ void stage_row_gray3(uint32_t* sm, const uint8_t* g, int wpix, int total_words, bool zero_row) {
    constexpr int L = kLead / 4;
    const int ngroups = (total_words - L + 2) / 3;          // 4 pixels -> 3 words
    for (int j = threadIdx.x; j < L; j += kThreads) sm[j] = 0u;
    for (int q = threadIdx.x; q < ngroups; q += kThreads) {
        uint32_t a = 0, b = 0, c = 0, d = 0;
        const int x = 4 * q;
        if (!zero_row) {
            if (x < wpix) a = __ldg(g + x);
            if (x + 1 < wpix) b = __ldg(g + x + 1);
            if (x + 2 < wpix) c = __ldg(g + x + 2);
            if (x + 3 < wpix) d = __ldg(g + x + 3);
        }
        const int w = L + 3 * q;
        if (w < total_words) sm[w] = a * 0x010101u | (b << 24);
        if (w + 1 < total_words) sm[w + 1] = b * 0x0101u | (c * 0x0101u << 16);
        if (w + 2 < total_words) sm[w + 2] = c | (d * 0x010101u << 8);
    }
}

// The five words a thread needs for 16 row bytes starting at row byte i (multiple of 16): rs[j] = row bytes
// [i + 4(j-1), i + 4j) for j = 0..4, i.e. rs[0] is the 4 bytes in front of the chunk.
__device__ __forceinline__ void load_chunk(const uint32_t* sm, int o, int i, uint32_t rs[5]) {
    const int w = (kLead + i) >> 2;                         // word of row byte i when o == 0; multiple of 4
    const uint32_t m1 = sm[w - 1];
    const uint4 q = *reinterpret_cast<const uint4*>(sm + w);
    const uint32_t p4 = sm[w + 4];
    const int sh = 8 * o;
    rs[0] = __funnelshift_r(m1, q.x, sh); rs[1] = __funnelshift_r(q.x, q.y, sh); rs[2] = __funnelshift_r(q.y, q.z, sh);
    rs[3] = __funnelshift_r(q.z, q.w, sh); rs[4] = __funnelshift_r(q.w, p4, sh);
}

}  // namespace

// dynamic smem: 3 row buffers (prev / cur rotate, out) of `words` words each; words = (max_rowbytes + 64) / 4 rounded to 4
__global__ void __launch_bounds__(kThreads) k_png_filter(const PageD* __restrict__ pages, int optimize,
                                                         uint32_t* __restrict__ row_adler, uint8_t* __restrict__ row_busy, int words) {
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ uint32_t red[5][kThreads / 32];
    __shared__ uint32_t red2[2][kThreads / 32];
    const PageD& P = pages[blockIdx.y];
    const int y0 = blockIdx.x * kRowsPerCta;
    if (y0 >= P.h) return;
    const int y1 = min(P.h, y0 + kRowsPerCta);
    const int bpp = P.c;
    const bool gray3 = P.pc == 1 && P.c == 3;                 // single-channel pixels, RGB output: replicate while staging
    const int n = P.w * bpp;                                  // row bytes
    const int64_t L = (int64_t)n + 1;                         // output row bytes
    const int nchunks = (n + 15) >> 4;
    const int lsh = 8 * (4 - bpp);                            // left = funnelshift(previous 4 bytes, these 4 bytes, lsh)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* bufA = smem; uint32_t* bufB = smem + words; uint32_t* smo = smem + 2 * words;
    // previous row of the first row
    {
        const uint8_t* gprev = P.pix + (int64_t)(y0 - 1) * P.pix_stride;
        if (gray3) stage_row_gray3(bufA, gprev, P.w, words, y0 == 0); else stage_row(bufA, gprev, n, words, y0 == 0);
    }
    uint32_t* smp = bufA; uint32_t* smc = bufB;
    for (int y = y0; y < y1; y++) {
        const uint8_t* grow = P.pix + (int64_t)y * P.pix_stride;
        const int oc = gray3 ? 0 : (int)((uintptr_t)grow & 3);
        const int op = (y == 0 || gray3) ? 0 : (int)((uintptr_t)(grow - P.pix_stride) & 3);
        if (gray3) stage_row_gray3(smc, grow, P.w, words, false); else stage_row(smc, grow, n, words, false);
        __syncthreads();

        // ---- is the row identical to the one above (blank paper)? then Up (or None for an all-zero row) with zero residuals
        uint32_t diff = 0, nonzero = 0;
        for (int ch = threadIdx.x; ch < nchunks; ch += kThreads) {
            uint32_t c[5], p[5];
            load_chunk(smc, oc, 16 * ch, c); load_chunk(smp, op, 16 * ch, p);
#pragma unroll
            for (int j = 1; j < 5; j++) { diff |= c[j] ^ p[j]; nonzero |= c[j]; }   // bytes past the row are zero in both
        }
        const int any_diff = __syncthreads_or((int)diff);
        if (threadIdx.x == 0 && row_busy && !any_diff) row_busy[P.row0 + y] = 0;          // cost hint for the LZ work queue
        int ftype;
        if (!any_diff) {
            const int any_nz = __syncthreads_or((int)nonzero);
            ftype = any_nz ? 2 : 0;
            for (int j = threadIdx.x; j < words; j += kThreads) smo[j] = 0u;
            __syncthreads();
            if (threadIdx.x == 0) {
                smo[kLead / 4 - 1] = (uint32_t)ftype << 24;
                row_adler[P.row0 + y] = ((uint32_t)(((uint64_t)L * ftype) % 65521u) << 16) | (uint32_t)ftype;
            }
        } else {
            // ---- pass 1: the candidate sums
            uint32_t s_none = 0, s_up = 0, s_sub = 0, s_avg = 0, s_pae = 0;
            for (int ch = threadIdx.x; ch < nchunks; ch += kThreads) {
                uint32_t c[5], p[5];
                load_chunk(smc, oc, 16 * ch, c); load_chunk(smp, op, 16 * ch, p);
#pragma unroll
                for (int j = 1; j < 5; j++) {
                    const uint32_t cur = c[j], up = p[j];
                    const uint32_t left = __funnelshift_r(c[j - 1], c[j], lsh), ul = __funnelshift_r(p[j - 1], p[j], lsh);
                    uint32_t mask = 0xFFFFFFFFu;
                    const int rem = n - (16 * ch + 4 * (j - 1));
                    if (rem < 4) mask = rem <= 0 ? 0u : (0xFFFFFFFFu >> (8 * (4 - rem)));
                    s_none = score_acc(cur & mask, s_none);
                    s_up = score_acc(__vsub4(cur, up) & mask, s_up);
                    s_sub = score_acc(__vsub4(cur, left) & mask, s_sub);
                    if (optimize) s_avg = score_acc(__vsub4(cur, __vhaddu4(left, up)) & mask, s_avg);
                    s_pae = score_acc(__vsub4(cur, paeth4(left, up, ul)) & mask, s_pae);
                }
            }
            s_none = warp_sum(s_none); s_up = warp_sum(s_up); s_sub = warp_sum(s_sub); s_avg = warp_sum(s_avg); s_pae = warp_sum(s_pae);
            if (lane == 0) { red[0][warp] = s_none; red[1][warp] = s_up; red[2][warp] = s_sub; red[3][warp] = s_avg; red[4][warp] = s_pae; }
            __syncthreads();
            uint32_t tot[5];
#pragma unroll
            for (int k = 0; k < 5; k++) { uint32_t t = 0;
#pragma unroll
                for (int w = 0; w < kThreads / 32; w++) t += red[k][w]; tot[k] = t; }
            // ZipEncode.c order: None is the incumbent; Up, Sub, (Avg), Paeth replace it only if strictly smaller,
            // and nothing is tried once the incumbent's sum is 0.
            ftype = 0; uint32_t best = tot[0];
            if (best > 0 && tot[1] < best) { best = tot[1]; ftype = 2; }
            if (best > 0 && tot[2] < best) { best = tot[2]; ftype = 1; }
            if (optimize && best > 0 && tot[3] < best) { best = tot[3]; ftype = 3; }
            if (best > 0 && tot[4] < best) { best = tot[4]; ftype = 4; }
            if (threadIdx.x == 0 && row_busy) row_busy[P.row0 + y] = (uint8_t)(1u + min(254u, best >> 9));   // more ink, more LZ work

            // ---- pass 2: materialise the winner as words (smo word kLead/4 + g = residual of row bytes 4g..4g+3,
            //      the word in front of it carries the filter byte in its top byte), Adler partials on the way
            uint32_t a_s1 = 0; uint64_t a_s2 = 0;             // sum b, sum (L-1-k)*b over this thread's bytes (k = index in out row)
            if (threadIdx.x == 0) { smo[kLead / 4 - 1] = (uint32_t)ftype << 24; a_s1 = ftype; a_s2 = (uint64_t)L * ftype; }
            for (int ch = threadIdx.x; ch < nchunks; ch += kThreads) {
                uint32_t c[5], p[5];
                load_chunk(smc, oc, 16 * ch, c); load_chunk(smp, op, 16 * ch, p);
                uint32_t r[4];
#pragma unroll
                for (int j = 1; j < 5; j++) {
                    const uint32_t cur = c[j];
                    uint32_t v;
                    if (ftype == 0) v = cur;
                    else if (ftype == 2) v = __vsub4(cur, p[j]);
                    else {
                        const uint32_t left = __funnelshift_r(c[j - 1], c[j], lsh);
                        if (ftype == 1) v = __vsub4(cur, left);
                        else if (ftype == 3) v = __vsub4(cur, __vhaddu4(left, p[j]));
                        else v = __vsub4(cur, paeth4(left, p[j], __funnelshift_r(p[j - 1], p[j], lsh)));
                    }
                    const int i = 16 * ch + 4 * (j - 1);
                    const int rem = n - i;
                    if (rem < 4) v &= rem <= 0 ? 0u : (0xFFFFFFFFu >> (8 * (4 - rem)));
                    r[j - 1] = v;
                    const uint32_t s = __dp4a(v, 0x01010101u, 0u);
                    a_s1 += s;
                    a_s2 += (uint64_t)(uint32_t)(L - 1 - i) * s - __dp4a(v, 0x03020100u, 0u);   // masked bytes are 0
                }
                *reinterpret_cast<uint4*>(smo + kLead / 4 + 4 * ch) = make_uint4(r[0], r[1], r[2], r[3]);
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
        }
        __syncthreads();
        // ---- copy out.  Output byte k (0 = filter byte) is shared byte kLead - 1 + k; the aligned global word m holds
        //      output bytes [4m - a0, 4m - a0 + 4) = funnelshift(smo[W + m], smo[W + m + 1], 8 * (3 - a0)), W = kLead/4 - 1.
        {
            uint8_t* gout = P.filt + (int64_t)y * L;
            const int a0 = (int)((uintptr_t)gout & 3);
            const int W = kLead / 4 - 1;
            const int sh = 8 * (3 - a0);
            const int first_w = a0 ? 1 : 0;                       // first fully covered word
            const int end_w = (int)((a0 + L) >> 2);               // one past the last fully covered word
            uint32_t* gw = (uint32_t*)(gout - a0);
            const uint8_t* so = (const uint8_t*)smo + kLead - 1;   // so[k] = output byte k
            for (int m = first_w + threadIdx.x; m < end_w; m += kThreads)
                gw[m] = __funnelshift_r(smo[W + m], smo[W + m + 1], sh);
            if (end_w > first_w) {
                if (threadIdx.x < 4) {
                    const int k = threadIdx.x;                    // head bytes: output offsets [0, 4 - a0)
                    if (a0 && k < 4 - a0) gout[k] = so[k];
                } else if (threadIdx.x < 8) {
                    const int64_t k = (int64_t)end_w * 4 - a0 + (threadIdx.x - 4);   // tail bytes
                    if (k < L) gout[k] = so[k];
                }
            } else {                                              // row shorter than one aligned word: plain bytes
                for (int k = threadIdx.x; k < L; k += kThreads) gout[k] = so[k];
            }
        }
        __syncthreads();                                          // smo / smp are reused by the next row
        uint32_t* t = smp; smp = smc; smc = t;
    }
}

int launch_png_filter(const PageD* d_pages, int npages, int max_h, int max_rowbytes, int optimize,
                         uint32_t* row_adler, uint8_t* row_busy, cudaStream_t st) {
    if (npages == 0 || max_h == 0) return 0;
    const int words = ((max_rowbytes + 64) / 4 + 3) & ~3;
    const size_t smem = (size_t)words * 3 * sizeof(uint32_t);
    if (smem > 48 * 1024) {
        // per-device function attribute (one process may drive several GPUs): remember what each device was given
        static std::atomic<size_t> configured_by_dev[kMaxDevices];
        int dev = 0;
        cudaGetDevice(&dev);
        const int slot = dev < 0 ? 0 : (dev < kMaxDevices ? dev : kMaxDevices - 1);
        if (dev >= kMaxDevices || smem > configured_by_dev[slot].load(std::memory_order_acquire)) {
            cudaFuncSetAttribute(k_png_filter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            configured_by_dev[slot].store(smem, std::memory_order_release);
        }
    }
    dim3 grid((max_h + kRowsPerCta - 1) / kRowsPerCta, npages);
    k_png_filter<<<grid, kThreads, smem, st>>>(d_pages, optimize, row_adler, row_busy, words);
    return 1;
}

// ---------------------------------------------------------------------------------------- Adler-32 combine
// One warp per page.  A row of L bytes with partials (s1, s2) acts on the running sums as
//   b += L*a + s2 ; a += s1   (mod 65521)      [zlib adler32.c semantics, a=1,b=0 at the start]
// Every lane folds a contiguous range of rows starting from (0,0); lane 0 then chains the 32 ranges with the same rule
// (a range of N bytes with sums (A,B) is just a long row).
__global__ void __launch_bounds__(128) k_adler_combine(const PageD* __restrict__ pages, int npages, const uint32_t* __restrict__ row_adler,
                                                       uint32_t* __restrict__ page_adler) {
    const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (p >= npages) return;
    const PageD& P = pages[p];
    const uint32_t L = (uint32_t)(((int64_t)P.w * P.c + 1) % 65521);
    const int rpl = (P.h + 31) / 32;
    const int r0 = min(P.h, lane * rpl), r1 = min(P.h, r0 + rpl);
    uint32_t A = 0, Bs = 0;
    for (int y = r0; y < r1; y++) {
        const uint32_t v = __ldg(row_adler + P.row0 + y);
        Bs = (uint32_t)((Bs + (uint64_t)L * A + (v >> 16)) % 65521u);
        A = (A + (v & 0xFFFFu)) % 65521u;
    }
    const uint32_t N = (uint32_t)(((uint64_t)(r1 - r0) * L) % 65521u);
    uint32_t a = 1, b = 0;
    for (int l = 0; l < 32; l++) {
        const uint32_t Al = __shfl_sync(0xffffffffu, A, l), Bl = __shfl_sync(0xffffffffu, Bs, l), Nl = __shfl_sync(0xffffffffu, N, l);
        b = (uint32_t)((b + (uint64_t)Nl * a + Bl) % 65521u);
        a = (a + Al) % 65521u;
    }
    if (lane == 0) page_adler[p] = (b << 16) | a;
}

int launch_adler_combine(const PageD* d_pages, int npages, const uint32_t* row_adler, uint32_t* page_adler, cudaStream_t st) {
    if (npages == 0) return 0;
    k_adler_combine<<<(npages + 3) / 4, 128, 0, st>>>(d_pages, npages, row_adler, page_adler);
    return 1;
}

}  // namespace vcp
