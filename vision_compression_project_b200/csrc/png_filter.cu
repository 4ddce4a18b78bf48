// png_filter.cu — PNG adaptive row filtering, bit-exact with Pillow's ZipEncode.c rule, plus the
// Adler-32 partials of the filtered stream.
//
// Replaces the per-row part of `page_image.save(path)` (backend/app/pipeline/pdf_extract.py:130):
// Pack.c (RGBX -> packed row) + ZipEncode.c (candidates None/Up/Sub/[Avg]/Paeth, pick the smallest
// sum of |signed residual| in the order Up, Sub, Avg(optimize only), Paeth with a strict '<' and an
// early-out once a sum is 0).  Restated in oracle/restate.py:filter_row and pinned by the 23 recorded
// PNGs of the reference (tests/golden/fixtures.json: 2339 filter decisions per page).
//
// One CTA walks 16 consecutive rows of a page, so every row is read from global memory once (the previous row stays in shared
// memory).  Rows are fetched by TMA bulk copies into a ring of row buffers, two rows ahead of the one being filtered: an elected
// thread arms the buffer's mbarrier with the byte count and issues one cp.async.bulk from the row's address rounded down to 16
// bytes — page rows of 3-byte pixels share no word phase, the copy engine does not care — so the row lands with its own byte offset
// o = address & 15, and no thread spends issue slots on staging.  A thread owns 16 contiguous shared-memory bytes of the current
// row (one 128-bit load + the word in front of them) and reads the bytes above them from the previous row's buffer at the
// relative offset of the two rows (six word loads, funnel-shifted).  Every candidate is evaluated 4 bytes per instruction:
// |cur - predictor| by VABSDIFF4, and the sum of |signed residual| = 128 - ||cur - predictor| - 128| by one accumulating
// VABSDIFF4 against 0x80808080 (the modular residual itself is only formed for the winner); Paeth is byte-SIMD emulation, its
// predictor is kept in shared memory for the second pass.  Sums are reduced with warp shuffles; the winning residual is
// materialised as whole words in the layout of the input row and leaves as aligned 32-bit stores (the filter byte makes the output
// row 1 + W*bpp long, so it is never aligned with its input).  A row identical to the one above (blank paper) skips the candidates
// and is written as zeros straight from registers.  HBM-bound by design: read W*bpp + write 1 + W*bpp per row.
#include "vcp_internal.cuh"
#include <atomic>

namespace vcp {

namespace {

#ifndef VCP_FILTER_THREADS
#define VCP_FILTER_THREADS 128
#endif
constexpr int kThreads = VCP_FILTER_THREADS;
constexpr int kRowsPerCta = 16;
constexpr int kLead = 32;        // bytes in front of the TMA destination of a row buffer: room for the pixel left of the first one
constexpr int kMaxRing = 3;

__device__ __forceinline__ uint32_t paeth4(uint32_t a, uint32_t b, uint32_t c) {
    const uint32_t pa = __vabsdiffu4(b, c), pb = __vabsdiffu4(a, c);
    const uint32_t same = ~(__vcmpgeu4(a, c) ^ __vcmpgeu4(b, c));          // (a-c) and (b-c) have the same sign
    const uint32_t pc = (same & __vaddus4(pa, pb)) | (~same & __vabsdiffu4(pa, pb));
    const uint32_t sa = __vcmpleu4(pa, pb) & __vcmpleu4(pa, pc);
    const uint32_t sb = ~sa & __vcmpleu4(pb, pc);
    return (a & sa) | (b & sb) | (c & ~(sa | sb));
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Gray -> RGB on the way in (Convert.c L/LA -> RGB is plain replication): `g` is a row of wpix gray bytes, the staged row is
// the 3*wpix interleaved bytes Pillow would have produced, at shared byte kLead (byte offset 0); zeros in front and behind.
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

// The words of shared-memory chunk ch of the current row (bytes [kLead + 16 ch - 4, kLead + 16 ch + 16) of its buffer): c[0] is
// the word in front of the chunk, c[1..4] the chunk.
__device__ __forceinline__ void load_cur(const uint32_t* buf, int ch, uint32_t c[5]) {
    const int w = (kLead >> 2) + 4 * ch;
    c[0] = buf[w - 1];
    const uint4 q = *reinterpret_cast<const uint4*>(buf + w);
    c[1] = q.x; c[2] = q.y; c[3] = q.z; c[4] = q.w;
}
// The same row bytes of the previous row, whose buffer holds them `delta` = (offset of prev) - (offset of cur) bytes further on.
__device__ __forceinline__ void load_prev(const uint32_t* buf, int ch, int delta, uint32_t p[5]) {
    const int a = kLead + 16 * ch + delta - 4;              // >= 13
    const uint32_t* w = buf + (a >> 2);
    const int sh = (a & 3) * 8;
    const uint32_t x0 = w[0], x1 = w[1], x2 = w[2], x3 = w[3], x4 = w[4], x5 = w[5];
    p[0] = __funnelshift_r(x0, x1, sh); p[1] = __funnelshift_r(x1, x2, sh); p[2] = __funnelshift_r(x2, x3, sh);
    p[3] = __funnelshift_r(x3, x4, sh); p[4] = __funnelshift_r(x4, x5, sh);
}
// byte mask of word j (1..4) of chunk ch: row byte indices [i0, i0 + 4) that lie inside [0, n)
__device__ __forceinline__ uint32_t word_mask(int i0, int n) {
    uint32_t m = 0xFFFFFFFFu;
    if (i0 < 0) m = i0 <= -4 ? 0u : m << (8 * -i0);
    const int rem = n - i0;
    if (rem < 4) m &= rem <= 0 ? 0u : 0xFFFFFFFFu >> (8 * (4 - rem));
    return m;
}

}  // namespace

// dynamic smem: `ring` input row buffers, the output row, the Paeth predictor row (if `cache`), `words` words each
__global__ void __launch_bounds__(kThreads) k_png_filter(const PageD* __restrict__ pages, int optimize,
                                                         uint32_t* __restrict__ row_adler, uint8_t* __restrict__ row_busy,
                                                         int words, int ring, int cache) {
    extern __shared__ __align__(128) uint32_t smem[];
    __shared__ uint32_t red[5][kThreads / 32];
    __shared__ uint32_t red2[2][kThreads / 32];
    __shared__ __align__(8) uint64_t bars[kMaxRing];
    const PageD& P = pages[blockIdx.y];
    const int y0 = blockIdx.x * kRowsPerCta;
    if (y0 >= P.h) return;
    const int y1 = min(P.h, y0 + kRowsPerCta);
    const int bpp = P.c;
    const bool gray3 = P.pc == 1 && P.c == 3;                 // single-channel pixels, RGB output: replicate while staging
    const int n = P.w * bpp;                                  // row bytes
    const int64_t L = (int64_t)n + 1;                         // output row bytes
    const int lsh = 8 * (4 - bpp);                            // left = funnelshift(previous 4 bytes, these 4 bytes, lsh)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* smo = smem + ring * words;
    uint32_t* smq = smo + words;
    const uint8_t* pix = P.pix; const int64_t stride = P.pix_stride;
    auto slot_of = [&](int y) { int u = y - (y0 - 1); while (u >= ring) u -= ring; return u; };   // rows are at most 16 + ring apart: no division
    auto buf_of = [&](int y) { return smem + slot_of(y) * words; };                        // row y (y0 - 1 = the row above the first)
    auto off_of = [&](int y) { return (gray3 || y < 0) ? 0 : (int)((uintptr_t)(pix + (int64_t)y * stride) & 15); };
    auto fetch = [&](int y) {                                 // one thread: row y -> its ring slot
        const int slot = slot_of(y);
        const uintptr_t a = (uintptr_t)(pix + (int64_t)y * stride);
        const uint32_t bytes = (uint32_t)((((a & 15) + n) + 15) & ~15);
        mbar_expect_tx(&bars[slot], bytes);
        bulk_g2s(reinterpret_cast<uint8_t*>(smem + slot * words) + kLead, reinterpret_cast<const void*>(a & ~(uintptr_t)15), bytes, &bars[slot]);
    };
    // ---- set-up: the lead bytes of every ring slot are zero for good (the copies land behind them); the row above the first
    for (int j = threadIdx.x; j < ring * (kLead / 4); j += kThreads) smem[(j / (kLead / 4)) * words + j % (kLead / 4)] = 0u;
    if (y0 == 0 && !gray3) for (int j = threadIdx.x; j < words; j += kThreads) buf_of(-1)[j] = 0u;
    if (!gray3 && threadIdx.x == 0) {
        for (int k = 0; k < ring; k++) mbar_init(&bars[k], 1);
    }
    __syncthreads();
    if (!gray3 && threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (y0 > 0) fetch(y0 - 1);
        else mbar_expect_tx(&bars[0], 0);                    // nothing to fetch above the first row: complete the slot's first phase empty
        for (int y = y0; y < min(y1, y0 + ring - 1); y++) fetch(y);
    }
    if (gray3) stage_row_gray3(buf_of(y0 - 1), pix + (int64_t)(y0 - 1) * stride, P.w, words, y0 == 0);
    if (!gray3 && y0 > 0) {
        mbar_wait(&bars[0], 0);
        const int o = off_of(y0 - 1);
        if (threadIdx.x < 4) reinterpret_cast<uint8_t*>(buf_of(y0 - 1))[kLead + o - 4 + threadIdx.x] = 0;
    }
    for (int y = y0; y < y1; y++) {
        uint32_t* smc = buf_of(y); const uint32_t* smp = buf_of(y - 1);
        const int oc = off_of(y), delta = off_of(y - 1) - oc;
        const int nchunks = (oc + n + 15) >> 4;
        if (gray3) stage_row_gray3(smc, pix + (int64_t)y * stride, P.w, words, false);
        else {
            const int use = y - (y0 - 1);
            int slot = use, round = 0;
            while (slot >= ring) { slot -= ring; round ^= 1; }
            mbar_wait(&bars[slot], round);
            if (threadIdx.x < 4) reinterpret_cast<uint8_t*>(smc)[kLead + oc - 4 + threadIdx.x] = 0;   // left of the first pixel
        }
        __syncthreads();

        // ---- is the row identical to the one above (blank paper)? then Up (or None for an all-zero row) with zero residuals
        uint32_t diff = 0, nonzero = 0;
        for (int ch = threadIdx.x; ch < nchunks; ch += kThreads) {
            uint32_t c[5], p[5];
            load_cur(smc, ch, c); load_prev(smp, ch, delta, p);
            const int i0 = 16 * ch - oc;
            if (i0 >= 0 && i0 + 16 <= n) {
#pragma unroll
                for (int j = 1; j < 5; j++) { diff |= c[j] ^ p[j]; nonzero |= c[j]; }
            } else {
#pragma unroll
                for (int j = 1; j < 5; j++) { const uint32_t m = word_mask(i0 + 4 * (j - 1), n); diff |= (c[j] ^ p[j]) & m; nonzero |= c[j] & m; }
            }
        }
        const int any_diff = __syncthreads_or((int)diff);
        uint8_t* gout = P.filt + (int64_t)y * L;
        const int a0 = (int)((uintptr_t)gout & 3);
        uint32_t* gw = (uint32_t*)(gout - a0);
        const int first_w = a0 ? 1 : 0;                       // first fully covered aligned output word
        const int end_w = (int)((a0 + L) >> 2);               // one past the last fully covered word
        int ftype;
        if (!any_diff) {
            const int any_nz = __syncthreads_or((int)nonzero);
            ftype = any_nz ? 2 : 0;
            if (threadIdx.x == 0) {
                if (row_busy) row_busy[P.row0 + y] = 0;                                    // cost hint for the LZ work queue
                row_adler[P.row0 + y] = ((uint32_t)(((uint64_t)L * ftype) % 65521u) << 16) | (uint32_t)ftype;
            }
            // the output row is the filter byte and zeros: straight from registers
            for (int m = first_w + threadIdx.x; m < end_w; m += kThreads) gw[m] = (m == 0) ? (uint32_t)ftype : 0u;
            if (end_w > first_w) {
                if (threadIdx.x < 4) { const int k = threadIdx.x; if (a0 && k < 4 - a0) gout[k] = k == 0 ? (uint8_t)ftype : (uint8_t)0; }
                else if (threadIdx.x < 8) { const int64_t k = (int64_t)end_w * 4 - a0 + (threadIdx.x - 4); if (k < L) gout[k] = k == 0 ? (uint8_t)ftype : (uint8_t)0; }
            } else {
                for (int k = threadIdx.x; k < L; k += kThreads) gout[k] = k == 0 ? (uint8_t)ftype : (uint8_t)0;
            }
        } else {
            // ---- pass 1: the candidate sums.  acc_x = sum over bytes of | |cur - pred| - 128 |; bytes outside the row count 128 each
            uint32_t s_none = 0, s_up = 0, s_sub = 0, s_avg = 0, s_pae = 0;
            for (int ch = threadIdx.x; ch < nchunks; ch += kThreads) {
                uint32_t c[5], p[5], pr[4];
                load_cur(smc, ch, c); load_prev(smp, ch, delta, p);
                const int i0 = 16 * ch - oc;
                const bool inner = i0 >= 0 && i0 + 16 <= n;
                // Flat paper: where the 20 bytes a chunk looks at are one value in both rows, every predictor is that value — all
                // residuals but None's are zero.  Most of a text row is such chunks (margins, gaps); a warp whose 32 chunks all
                // are skips the candidates together (no divergence), the predictor row still gets its (trivial) entry.
                const uint32_t v0 = c[1];
                const bool flat = inner && v0 == __byte_perm(v0, 0u, 0x0000) && c[0] == v0 && c[2] == v0 && c[3] == v0 && c[4] == v0 &&
                                  p[0] == v0 && p[1] == v0 && p[2] == v0 && p[3] == v0 && p[4] == v0;
                if (__all_sync(__activemask(), flat)) {
                    const uint32_t z = 4u * 512u;                              // four words of zero residuals: |0 - 128| per byte
                    s_none = 4u * __vsadu4(v0, 0x80808080u) + s_none;
                    s_up += z; s_sub += z; s_pae += z;
                    if (optimize) s_avg += z;
                    if (cache) *reinterpret_cast<uint4*>(smq + (kLead >> 2) + 4 * ch) = make_uint4(v0, v0, v0, v0);
                    continue;
                }
#pragma unroll
                for (int j = 1; j < 5; j++) {
                    const uint32_t cur = c[j], up = p[j];
                    const uint32_t left = __funnelshift_r(c[j - 1], c[j], lsh), ul = __funnelshift_r(p[j - 1], p[j], lsh);
                    const uint32_t m = inner ? 0xFFFFFFFFu : word_mask(i0 + 4 * (j - 1), n);
                    const uint32_t pae = paeth4(left, up, ul);
                    pr[j - 1] = pae;
                    s_none = __vsadu4(cur & m, 0x80808080u) + s_none;
                    s_up = __vsadu4(__vabsdiffu4(cur, up) & m, 0x80808080u) + s_up;
                    s_sub = __vsadu4(__vabsdiffu4(cur, left) & m, 0x80808080u) + s_sub;
                    if (optimize) s_avg = __vsadu4(__vabsdiffu4(cur, __vhaddu4(left, up)) & m, 0x80808080u) + s_avg;
                    s_pae = __vsadu4(__vabsdiffu4(cur, pae) & m, 0x80808080u) + s_pae;
                }
                if (cache) *reinterpret_cast<uint4*>(smq + (kLead >> 2) + 4 * ch) = make_uint4(pr[0], pr[1], pr[2], pr[3]);
            }
            s_none = warp_sum(s_none); s_up = warp_sum(s_up); s_sub = warp_sum(s_sub); s_avg = warp_sum(s_avg); s_pae = warp_sum(s_pae);
            if (lane == 0) { red[0][warp] = s_none; red[1][warp] = s_up; red[2][warp] = s_sub; red[3][warp] = s_avg; red[4][warp] = s_pae; }
            __syncthreads();
            uint32_t tot[5];
            const uint32_t full = 128u * 16u * (uint32_t)nchunks;
#pragma unroll
            for (int k = 0; k < 5; k++) { uint32_t t = 0;
#pragma unroll
                for (int w = 0; w < kThreads / 32; w++) t += red[k][w]; tot[k] = full - t; }
            // ZipEncode.c order: None is the incumbent; Up, Sub, (Avg), Paeth replace it only if strictly smaller,
            // and nothing is tried once the incumbent's sum is 0.
            ftype = 0; uint32_t best = tot[0];
            if (best > 0 && tot[1] < best) { best = tot[1]; ftype = 2; }
            if (best > 0 && tot[2] < best) { best = tot[2]; ftype = 1; }
            if (optimize && best > 0 && tot[3] < best) { best = tot[3]; ftype = 3; }
            if (best > 0 && tot[4] < best) { best = tot[4]; ftype = 4; }
            if (threadIdx.x == 0 && row_busy) row_busy[P.row0 + y] = (uint8_t)(1u + min(254u, best >> 9));   // more ink, more LZ work

            // ---- pass 2: materialise the winner in the layout of the input row (residual of row byte i at byte kLead + oc + i
            //      of smo, the filter byte in front of it), Adler partials on the way
            uint32_t a_s1 = 0; uint64_t a_s2 = 0;             // sum b, sum (L-1-k)*b over this thread's bytes (k = index in out row)
            for (int ch = threadIdx.x; ch < nchunks; ch += kThreads) {
                uint32_t c[5], p[5];
                load_cur(smc, ch, c);
                if (ftype >= 2) load_prev(smp, ch, delta, p);
                uint4 pq = make_uint4(0, 0, 0, 0);
                if (ftype == 4 && cache) pq = *reinterpret_cast<const uint4*>(smq + (kLead >> 2) + 4 * ch);
                const uint32_t pc4[4] = {pq.x, pq.y, pq.z, pq.w};
                const int i0 = 16 * ch - oc;
                const bool inner = i0 >= 0 && i0 + 16 <= n;
                uint32_t r[4];
                uint32_t c_s1 = 0, c_s2 = 0, c_sw = 0;               // this chunk: sum b, sum (j-1) * (word sum), sum of in-word weights
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
                        else v = __vsub4(cur, cache ? pc4[j - 1] : paeth4(left, p[j], __funnelshift_r(p[j - 1], p[j], lsh)));
                    }
                    const int i = i0 + 4 * (j - 1);
                    if (!inner) v &= word_mask(i, n);
                    r[j - 1] = v;
                    const uint32_t s = __dp4a(v, 0x01010101u, 0u);
                    c_s1 += s; c_s2 += (uint32_t)(j - 1) * s; c_sw = __dp4a(v, 0x03020100u, c_sw);   // masked bytes are 0
                }
                // byte b of word j-1 sits at row byte i0 + 4(j-1) + b and weighs L - 1 - that
                a_s1 += c_s1;
                a_s2 += (uint64_t)(int64_t)(L - 1 - i0) * c_s1 - (uint64_t)(4u * c_s2 + c_sw);
                *reinterpret_cast<uint4*>(smo + (kLead >> 2) + 4 * ch) = make_uint4(r[0], r[1], r[2], r[3]);
            }
            if (threadIdx.x == 0) { a_s1 += ftype; a_s2 += (uint64_t)L * ftype; }
            a_s1 %= 65521u;
            uint32_t a_s2m = (uint32_t)(a_s2 % 65521u);
            a_s1 = warp_sum(a_s1); a_s2m = warp_sum(a_s2m);
            if (lane == 0) { red2[0][warp] = a_s1; red2[1][warp] = a_s2m; }
            if (threadIdx.x == 0) reinterpret_cast<uint8_t*>(smo)[kLead + oc - 1] = (uint8_t)ftype;   // behind this thread's own chunk-0 store
            __syncthreads();
            if (threadIdx.x == 0) {
                uint32_t t1 = 0, t2 = 0;
                for (int w = 0; w < kThreads / 32; w++) { t1 += red2[0][w]; t2 += red2[1][w]; }
                row_adler[P.row0 + y] = ((t2 % 65521u) << 16) | (t1 % 65521u);
            }
            // ---- copy out.  Output byte k (0 = filter byte) is smo byte S0 + k, S0 = kLead + oc - 1; the aligned global word m holds
            //      output bytes [4m - a0, 4m - a0 + 4) = smo bytes from S0 + 4m - a0: two words, funnel-shifted.
            const int S0 = kLead + oc - 1;
            const int sh = ((S0 - a0) & 3) * 8;
            const uint32_t* sw = smo + ((S0 - a0) >> 2);
            const uint8_t* so = reinterpret_cast<const uint8_t*>(smo) + S0;   // so[k] = output byte k
            for (int m = first_w + threadIdx.x; m < end_w; m += kThreads) gw[m] = __funnelshift_r(sw[m], sw[m + 1], sh);
            if (end_w > first_w) {
                if (threadIdx.x < 4) { const int k = threadIdx.x; if (a0 && k < 4 - a0) gout[k] = so[k]; }                       // head bytes
                else if (threadIdx.x < 8) { const int64_t k = (int64_t)end_w * 4 - a0 + (threadIdx.x - 4); if (k < L) gout[k] = so[k]; }   // tail bytes
            } else {                                              // row shorter than one aligned word: plain bytes
                for (int k = threadIdx.x; k < L; k += kThreads) gout[k] = so[k];
            }
        }
        // ---- the slot of the row above is free now: fetch the row `ring - 1` ahead into it
        if (!gray3) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // this thread's writes into the slot (left-of-first bytes) before the copy engine's
        __syncthreads();
        if (!gray3 && threadIdx.x == 0 && y + ring - 1 < y1) fetch(y + ring - 1);
    }
}

int launch_png_filter(const PageD* d_pages, int npages, int max_h, int max_rowbytes, int optimize,
                         uint32_t* row_adler, uint8_t* row_busy, cudaStream_t st) {
    if (npages == 0 || max_h == 0) return 0;
    const int words = (((max_rowbytes + kLead + 16 + 96) / 4) + 3) & ~3;      // lead + offset + row + the words read past it
    // ring of 3 input rows + output row + Paeth row when that fits an SM's shared memory comfortably; very wide rows get 2 + 1
    int ring = 3, cache = 1;
    if ((size_t)words * 4 * 5 > 200 * 1024) { cache = 0; if ((size_t)words * 4 * 4 > 200 * 1024) ring = 2; }
    const size_t smem = (size_t)words * 4 * (ring + 1 + cache);
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
    k_png_filter<<<grid, kThreads, smem, st>>>(d_pages, optimize, row_adler, row_busy, words, ring, cache);
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
