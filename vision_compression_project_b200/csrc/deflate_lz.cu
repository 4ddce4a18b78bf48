// deflate_lz.cu — window-parallel LZ77 match finding for the zlib stream inside the PNG.
//
// Replaces zlib's deflate.c (longest_match / deflate_slow), which Pillow's ZipEncode.c drives when the
// reference calls page_image.save(...) (backend/app/pipeline/pdf_extract.py:130).  The output is NOT
// byte-identical to zlib's (any valid deflate stream decodes to the same filtered bytes; the contract is
// validity and size <= 1.05 x Pillow's default); tests compare the token stream
// bit-for-bit with the sequential model in tests/model/deflate_model.c and inflate the result with zlib.
//
// Decomposition: the filtered stream of a page is cut into 32 KiB sub-chunks, one WARP each.  A warp
// owns two small hash tables in shared memory (3-byte hash: 2^11 buckets x 2 ways, 6-byte hash: 2^10 x 2,
// u16 positions relative to sub-chunk start - 32 KiB, so the previous 32 KiB of the page are addressable
// history and are inserted before the sub-chunk starts).  It then walks its sub-chunk in windows of 32
// positions, one per lane:
//   1. one coalesced 128-byte load of the window, bytes handed to lanes with shuffles;
//   2. every lane proposes a match: distance-1 and distance-bpp runs from two 64-bit ballot masks,
//      up to four hash candidates verified against global memory (exact up to 64 bytes);
//      short matches are priced against literals with the running histogram (quarter-bit log2);
//   3. one-step lazy rule between neighbouring lanes (a shuffle), greedy parse from lane 0 resolved by
//      5 rounds of pointer jumping, last token extended cooperatively to <= 258, maximal distance-1
//      runs continued without re-hashing;
//   4. tokens written compactly (rank = popc of the selection mask), histogram by shared atomics,
//      window positions inserted with atomicMax (highest position wins -> deterministic).
// Integer/latency bound, not HBM bound: the stream is read ~once from L2/HBM; see DESIGN.md §5.
#include "vcp_internal.cuh"

namespace vcp {

namespace {

constexpr int HB3 = 11;               // 3-byte-hash table: 2^11 buckets, u32 = (newest<<16) | older
constexpr int HB6 = 10;               // 6-byte-hash table: 2^10 buckets, same bucket format
constexpr int kLzWarps = 4;           // warps (= sub-chunks) per CTA
constexpr int kLaneCap = 64;          // exact compare length of a hash candidate inside a lane
constexpr int kLazyMax = 16;
constexpr int kCostMaxLen = 8;
constexpr int kCostWarm = 64;

struct __align__(16) WarpMem {
    uint32_t t3[1 << HB3];
    uint32_t t6[1 << HB6];
    uint32_t hist[320];
};

__device__ __forceinline__ uint32_t ldu(const uint32_t* __restrict__ S32, int x) {   // unaligned u32 at byte x
    const int w = x >> 2;
    return __funnelshift_r(__ldg(S32 + w), __ldg(S32 + w + 1), (x & 3) * 8);
}

__device__ __forceinline__ int ilog2x4(uint32_t v) {   // quarter-bit log2, v >= 1
    const int n = 31 - __clz(v);
    const uint32_t frac = n >= 2 ? (v >> (n - 2)) & 3u : (n == 1 ? (v & 1u) << 1 : 0u);
    return 4 * n + (int)frac;
}

__device__ __forceinline__ int len_sym(int len) {      // 3..258 -> 0..28
    const int v = len - 3;
    if (v < 8) return v;
    if (len == 258) return 28;
    const int n = 31 - __clz(v);
    return 4 * (n - 1) + ((v >> (n - 2)) & 3);
}
__device__ __forceinline__ int dist_sym(int dist) {    // 1..32768 -> 0..29
    const int v = dist - 1;
    if (v < 4) return v;
    const int n = 31 - __clz(v);
    return 2 * n + ((v >> (n - 1)) & 1);
}

// exact match length of S[q..] vs S[c..], up to cap bytes (first 4 bytes of q given)
__device__ __forceinline__ int lane_match(const uint32_t* __restrict__ S32, int q, int c, uint32_t cur4, int cap) {
    int wc = c >> 2; const int shc = (c & 3) * 8;
    uint32_t c0 = __ldg(S32 + wc), c1 = __ldg(S32 + wc + 1);
    uint32_t x = cur4 ^ __funnelshift_r(c0, c1, shc);
    if (x) return min(((__ffs(x) - 1) >> 3), cap);
    int wq = (q >> 2) + 1; const int shq = (q & 3) * 8;
    uint32_t q0 = __ldg(S32 + wq);
    int n = 4;
    while (n < cap) {
        const uint32_t q1 = __ldg(S32 + wq + 1);
        c0 = c1; c1 = __ldg(S32 + wc + 2);
        x = __funnelshift_r(q0, q1, shq) ^ __funnelshift_r(c0, c1, shc);
        if (x) { n += (__ffs(x) - 1) >> 3; break; }
        q0 = q1; wq++; wc++; n += 4;
    }
    return min(n, cap);
}

// warp-cooperative: length of the common prefix of S[y..] and S[y-d..], up to maxn (128 bytes per step)
__device__ __forceinline__ int coop_match(const uint32_t* __restrict__ S32, int y, int d, int maxn, int lane) {
    int n = 0;
    while (n < maxn) {
        const int k = y + n + 4 * lane;
        const uint32_t x = ldu(S32, k) ^ ldu(S32, k - d);
        const uint32_t mism = __ballot_sync(0xffffffffu, x != 0);
        if (mism) {
            const int first = __ffs(mism) - 1;
            const uint32_t xx = __shfl_sync(0xffffffffu, x, first);
            n += 4 * first + ((__ffs(xx) - 1) >> 3);
            break;
        }
        n += 128;
    }
    return min(n, maxn);
}

}  // namespace

__global__ void __launch_bounds__(kLzWarps * 32) k_lz(BatchD B) {
    extern __shared__ __align__(16) unsigned char lz_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = blockIdx.x * kLzWarps + warp;
    if (sub >= B.nsub) return;
    WarpMem& M = reinterpret_cast<WarpMem*>(lz_smem)[warp];
    const BlockD& blk = B.blocks[B.sub2blk[sub]];
    const PageD& pg = B.pages[blk.page];
    const uint8_t* __restrict__ S = pg.filt;
    const uint32_t* __restrict__ S32 = reinterpret_cast<const uint32_t*>(S);     // page streams are 256-byte aligned
    const int F = (int)pg.filt_len;
    const int bpp = pg.c;
    const int s = (int)blk.start + (sub - blk.sub0) * kSubBytes;
    const int e = min(s + kSubBytes, (int)(blk.start + blk.len));
    const int base = s - kMaxDist;                                               // table entries are pos - base (u16), 0 = empty
    uint32_t* __restrict__ tok = B.tokens + ((S - B.filt_base) + s);

    // ---- clear tables + histogram
    {
        uint4* z = reinterpret_cast<uint4*>(&M);
        const uint4 zero = make_uint4(0, 0, 0, 0);
        for (int i = lane; i < (int)(sizeof(WarpMem) / 16); i += 32) z[i] = zero;
    }
    __syncwarp();

    // ---- prime with the previous 32 KiB of the page (window by window, same arithmetic as the main loop's insert)
    for (int w0 = max(0, s - kMaxDist); w0 < s; w0 += 32) {
        const int A = w0 & ~3;
        const uint32_t wv = __ldg(S32 + (A >> 2) + lane);
        const int o = (w0 - A) + lane, k = o >> 2, sh = (o & 3) * 8;
        const uint32_t a0 = __shfl_sync(0xffffffffu, wv, k), a1 = __shfl_sync(0xffffffffu, wv, k + 1), a2 = __shfl_sync(0xffffffffu, wv, k + 2);
        const uint32_t cur4 = __funnelshift_r(a0, a1, sh), nxt4 = __funnelshift_r(a1, a2, sh);
        const int q = w0 + lane;
        const bool ok3 = q < s && q + 2 < F, ok6 = q < s && q + 6 <= F;
        const uint32_t h3 = ((cur4 & 0xFFFFFFu) * 0x9E3779B1u) >> (32 - HB3);
        const uint32_t h6 = (cur4 * 0x9E3779B1u + (nxt4 & 0xFFFFu) * 0x85EBCA77u) >> (32 - HB6);
        const uint32_t b3 = M.t3[h3], b6 = M.t6[h6];
        __syncwarp();
        const uint32_t pos = (uint32_t)(q - base);
        if (ok3) atomicMax(&M.t3[h3], (pos << 16) | (b3 >> 16));
        if (ok6) atomicMax(&M.t6[h6], (pos << 16) | (b6 >> 16));
        __syncwarp();
    }

    // ---- main loop
    int p = s;
    uint32_t ntok = 0;
    while (p < e) {
        const int A = (p - 4) & ~3;                                              // >= -4: the pad in front of the stream is addressable
        const uint32_t wv = __ldg(S32 + (A >> 2) + lane);
        const int q = p + lane;
        // bytes [q-4, q+8) for the first half, [q+28, q+33) for the second half
        const int o = (p - 4 - A) + lane, k = o >> 2, sh = (o & 3) * 8;
        const uint32_t a0 = __shfl_sync(0xffffffffu, wv, k), a1 = __shfl_sync(0xffffffffu, wv, k + 1);
        const uint32_t a2 = __shfl_sync(0xffffffffu, wv, k + 2), a3 = __shfl_sync(0xffffffffu, wv, k + 3);
        const uint32_t lo = __funnelshift_r(a0, a1, sh), cur4 = __funnelshift_r(a1, a2, sh), nxt4 = __funnelshift_r(a2, a3, sh);
        const uint32_t c0 = __shfl_sync(0xffffffffu, wv, k + 8), c1 = __shfl_sync(0xffffffffu, wv, k + 9);
        const uint32_t lo2 = __funnelshift_r(c0, c1, sh);                        // bytes [q+28, q+32)
        const uint32_t b2 = (c1 >> sh) & 0xFFu;                                  // byte q+32
        // equality bits for distance 1 and distance bpp at positions q and q+32
        const uint32_t bq = cur4 & 0xFFu;
        const bool e1a = (bq == (lo >> 24)) && q >= 1;
        const bool eba = (bq == ((lo >> (8 * (4 - bpp))) & 0xFFu)) && q >= bpp;
        const bool e1b = (b2 == (lo2 >> 24));
        const bool ebb = (b2 == ((lo2 >> (8 * (4 - bpp))) & 0xFFu));
        const unsigned long long m1 = (unsigned long long)__ballot_sync(0xffffffffu, e1a) | ((unsigned long long)__ballot_sync(0xffffffffu, e1b) << 32);
        const unsigned long long mb = (unsigned long long)__ballot_sync(0xffffffffu, eba) | ((unsigned long long)__ballot_sync(0xffffffffu, ebb) << 32);

        const int limit = min(kMaxMatch, e - q);                                 // <= 0 for lanes past the sub-chunk
        int bl = 0, bd = 0, be = 0; bool bcap = false;
        const uint32_t h3 = ((cur4 & 0xFFFFFFu) * 0x9E3779B1u) >> (32 - HB3);
        const uint32_t h6 = (cur4 * 0x9E3779B1u + (nxt4 & 0xFFFFu) * 0x85EBCA77u) >> (32 - HB6);
        const uint32_t b3 = M.t3[h3], b6 = M.t6[h6];
        if (limit >= 3) {
            const int runcap = min(64 - lane, limit);
            {   // distance 1
                const unsigned long long inv = ~(m1 >> lane);
                const int r = inv ? __ffsll((long long)inv) - 1 : 64;
                const int l = min(r, runcap);
                const bool c = (l == runcap) && (runcap < limit);
                const int eff = c ? 1000 : l;
                if (l >= 3 && eff > be) { be = eff; bl = l; bd = 1; bcap = c; }
            }
            if (bpp > 1) {
                const unsigned long long inv = ~(mb >> lane);
                const int r = inv ? __ffsll((long long)inv) - 1 : 64;
                const int l = min(r, runcap);
                const bool c = (l == runcap) && (runcap < limit);
                const int eff = c ? 1000 : l;
                if (l >= 3 && eff > be) { be = eff; bl = l; bd = bpp; bcap = c; }
            }
            if (!bcap) {                                                          // a capped run outranks every hash candidate
                const int hcap = min(kLaneCap, limit);
                const bool ok6 = q + 6 <= F;
#pragma unroll
                for (int w = 0; w < 4; w++) {
                    const uint32_t cnd = (w == 0) ? (b3 >> 16) : (w == 1) ? (b3 & 0xFFFFu) : (w == 2) ? (b6 >> 16) : (b6 & 0xFFFFu);
                    if (cnd == 0 || (w >= 2 && !ok6)) continue;
                    const int cp = base + (int)cnd;
                    const int d = q - cp;
                    if (d <= 0 || d > kMaxDist) continue;
                    const int l = lane_match(S32, q, cp, cur4, hcap);
                    const bool c = (l == hcap) && (hcap < limit);
                    const int eff = c ? 1000 : l;
                    if (l >= 3 && (eff > be || (eff == be && d < bd))) { be = eff; bl = l; bd = d; bcap = c; }
                }
            }
            // price short matches against literals with the histogram as of the window start
            if (bl >= 3 && bl <= kCostMaxLen && ntok >= kCostWarm) {
                const int lgN = ilog2x4(ntok + 1);
                int lit = 0;
#pragma unroll
                for (int kk = 0; kk < kCostMaxLen; kk++) {
                    if (kk < bl) {
                        const uint32_t byte = (kk < 4 ? (cur4 >> (8 * kk)) : (nxt4 >> (8 * (kk - 4)))) & 0xFFu;
                        lit += lgN - ilog2x4(M.hist[byte] + 1);
                    }
                }
                const int ls = bl - 3, ds = dist_sym(bd);                         // bl <= 8: no length extra bits
                const int dx = ds < 4 ? 0 : (ds >> 1) - 1;
                const int mc = (lgN - ilog2x4(M.hist[257 + ls] + 1)) + (lgN - ilog2x4(M.hist[286 + ds] + 1)) + 4 * dx;
                if (mc >= lit) { bl = 0; bd = 0; bcap = false; }
            }
        }
        // ---- one-step lazy rule between neighbouring lanes
        {
            const int nxt = __shfl_down_sync(0xffffffffu, bl, 1);
            if (lane < 31 && bl >= 3 && bl < kLazyMax && nxt > bl) { bl = 0; bd = 0; bcap = false; }
        }
        // ---- greedy parse from lane 0 by pointer jumping
        const int nvalid = min(32, e - p);
        int J = lane + (bl ? bl : 1);
        if (J >= nvalid) J = 32;
        uint32_t Msel = 1u << lane;
#pragma unroll
        for (int r = 0; r < 5; r++) {
            const uint32_t Mj = __shfl_sync(0xffffffffu, Msel, J & 31);
            const int Jj = __shfl_sync(0xffffffffu, J, J & 31);
            if (J < 32) { Msel |= Mj; J = Jj; }
        }
        const uint32_t sel = __shfl_sync(0xffffffffu, Msel, 0);
        const int last = 31 - __clz(sel);
        // ---- the last token may be capped: extend it cooperatively
        int Ll = __shfl_sync(0xffffffffu, bl, last);
        const int dl = __shfl_sync(0xffffffffu, bd, last);
        const bool capl = __shfl_sync(0xffffffffu, (int)bcap, last) != 0;
        const int ql = p + last;
        if (capl) {
            const int lim = min(kMaxMatch, e - ql);
            Ll += coop_match(S32, ql + Ll, dl, lim - Ll, lane);
        }
        if (lane == last) bl = Ll;
        __syncwarp();                                                            // all histogram reads of this window are done
        // ---- emit tokens + histogram
        if ((sel >> lane) & 1u) {
            const int rank = __popc(sel & ((1u << lane) - 1u));
            if (bl) {
                tok[ntok + rank] = 0x80000000u | ((uint32_t)(bd - 1) << 8) | (uint32_t)(bl - 3);
                atomicAdd(&M.hist[257 + len_sym(bl)], 1u);
                atomicAdd(&M.hist[286 + dist_sym(bd)], 1u);
            } else {
                tok[ntok + rank] = bq;
                atomicAdd(&M.hist[bq], 1u);
            }
        }
        ntok += __popc(sel);
        int next = ql + (Ll ? Ll : 1);
        // ---- continuation of maximal distance-1 runs
        if (Ll == kMaxMatch && dl <= 1) {
            uint32_t extra = 0;
            while (next < e) {
                const int lim = min(kMaxMatch, e - next);
                if (lim < kMaxMatch) break;
                if (coop_match(S32, next, dl, kMaxMatch, lane) < kMaxMatch) break;
                if (lane == 0) tok[ntok + extra] = 0x80000000u | ((uint32_t)(dl - 1) << 8) | (uint32_t)(kMaxMatch - 3);
                extra++; next += kMaxMatch;
            }
            if (extra) {
                if (lane == 0) { atomicAdd(&M.hist[257 + 28], extra); atomicAdd(&M.hist[286 + dist_sym(dl)], extra); }
                ntok += extra;
            }
        }
        // ---- insert this window's positions (atomicMax: the highest position of a bucket group wins)
        {
            const uint32_t pos = (uint32_t)(q - base);
            if (q < e && q + 2 < F) atomicMax(&M.t3[h3], (pos << 16) | (b3 >> 16));
            if (q < e && q + 6 <= F) atomicMax(&M.t6[h6], (pos << 16) | (b6 >> 16));
        }
        __syncwarp();
        p = next;
    }
    // ---- results
    if (lane == 0) B.sub_ntok[sub] = ntok;
    uint32_t* hout = B.sub_hist + (size_t)sub * kHistSize;
    for (int i = lane; i < kHistSize; i += 32) hout[i] = M.hist[i];
}

int launch_lz(const BatchD& b, cudaStream_t st) {
    if (b.nsub == 0) return 0;
    const size_t smem = sizeof(WarpMem) * kLzWarps;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(k_lz, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    k_lz<<<(b.nsub + kLzWarps - 1) / kLzWarps, kLzWarps * 32, smem, st>>>(b);
    return 1;
}

}  // namespace vcp
