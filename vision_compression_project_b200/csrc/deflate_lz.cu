// deflate_lz.cu — window-parallel LZ77 match finding for the zlib stream inside the PNG.
//
// Replaces zlib's deflate.c (longest_match / deflate_slow), which Pillow's ZipEncode.c drives when the
// reference calls page_image.save(...) (backend/app/pipeline/pdf_extract.py:130).  The output is NOT
// byte-identical to zlib's (any valid deflate stream decodes to the same filtered bytes; the contract is
// validity and size <= 1.05 x Pillow's default); tests compare the token stream
// bit-for-bit with the sequential model in tests/model/deflate_model.c and inflate the result with zlib.
//
// Decomposition: the filtered stream of a page is cut into 32 KiB sub-chunks, one WARP each (persistent warps pull
// sub-chunks from a work queue in stream order).  A warp owns two small hash tables in shared memory (3-byte and
// 4-byte keys, 2^10 buckets x 2 ways each, u16 positions relative to sub-chunk start - 32 KiB, so the previous 32 KiB
// of the page are addressable history and are inserted before the sub-chunk starts; 8.6 KB per warp -> 26 warps/SM).
// It then walks its sub-chunk in windows of 32 positions, one per lane:
//   1. the stream is pulled through three 128-byte register chunks per warp (current, next, prefetched), so a
//      window's bytes come from shuffles and the global-load latency is hidden behind the previous windows;
//   2. every lane proposes a match: distance-1 and distance-bpp runs from two 64-bit ballot masks; four table
//      candidates and the byte one filtered row up, compared 16 bytes deep (first 4, then 12, the loads of a
//      stage issued together); fully compared candidates are ranked by one packed key, the ones that hit the
//      16-byte cap are only remembered; short matches are priced against literals with the running histogram;
//   3. one-step lazy rule between neighbouring lanes (a shuffle); greedy parse from lane 0 by 5 rounds of pointer
//      jumping; when the parse starts a token on a lane with capped candidates, all of them are measured to the
//      258-byte limit cooperatively (128 bytes of every candidate in one round trip), the best wins, and the parse
//      is redone from there (at most two or three times per window); a maximal distance-1 run is measured 1 KiB
//      per round trip and emitted as a burst of 258-tokens without re-hashing;
//   4. tokens written compactly (rank = popc of the selection mask), histogram by shared atomics,
//      window positions inserted with atomicMax (highest position wins -> deterministic).
// Integer-issue / latency bound, not HBM bound: the stream is read ~once from L2/HBM; see DESIGN.md §4.
#include "vcp_internal.cuh"
#include <algorithm>
#include <atomic>

namespace vcp {

namespace {

// Effort classes, selected by Pillow's compress_level (1-3 fast, 4-6 default, 7-9 best).  Each is one instantiation of k_lz.
//   HB3 / HB6   log2 buckets of the 3-byte / 4-byte hash tables (u32 bucket = (newest << 16) | older)
//   NOISY       literal EMA (8 x literals per window) from which only the newest way of each table is verified
//   LAZY        one-step lazy rule applies to matches shorter than this (0 = greedy)
//   PROBE       row-above candidate: 0 none, 1 only where a run starts, 2 everywhere
struct CfgFast    { static constexpr int HB3 = 9,  HB6 = 9,  NOISY = 0,          LAZY = 0,  PROBE = 0; };
#ifndef VCP_HB3
#define VCP_HB3 10
#endif
#ifndef VCP_HB6
#define VCP_HB6 10
#endif
struct CfgDefault { static constexpr int HB3 = VCP_HB3, HB6 = VCP_HB6, NOISY = 160, LAZY = 16, PROBE = 1; };
struct CfgBest    { static constexpr int HB3 = 11, HB6 = 12, NOISY = 0x7fffffff, LAZY = 16, PROBE = 2; };
constexpr int kH2Bytes = 4;           // bytes keyed by the second table
static_assert(kGroupSubs == 1 || kSubBytes == 32768, "the table rebase of grouped sub-chunks shifts by 32 KiB");
static_assert(kPrimeBytes % 512 == 0 && kPrimeBytes <= kMaxDist && kSubBytes % 512 == 0 && kSubBytes <= kMaxDist, "sub-chunk geometry");
constexpr int kLzWarps = 2;           // warps (= sub-chunks) per CTA
constexpr int kLaneCap = 16;          // compare depth of a hash candidate inside a lane (deeper only for tokens the parse selects)
constexpr int kCostMaxLen = 8;
constexpr int kCostWarm = 64;
constexpr int kCostEpochLog2 = 6;     // the cost table is refreshed at the first window after every 64 tokens (deflate_model.c cost_epoch)
constexpr unsigned kFull = 0xffffffffu;

template <class Cfg>
struct __align__(16) WarpMem {
    uint32_t t3[1 << Cfg::HB3];
    uint32_t t6[1 << Cfg::HB6];
    uint32_t hist[160];          // 320 token counters, two 16-bit halves per word (a sub-chunk has at most kSubBytes tokens)
    uint8_t clg[320];            // quarter-bit log2 of (count + 1) per token symbol, as of the last cost epoch (pricing of short matches)
};

static_assert(kSubBytes < 65536, "token counters are 16-bit");
template <class WM> __device__ __forceinline__ uint32_t hist_get(const WM& M, int i) { return (M.hist[i >> 1] >> ((i & 1) * 16)) & 0xFFFFu; }
template <class WM> __device__ __forceinline__ void hist_add(WM& M, int i, uint32_t v) { atomicAdd(&M.hist[i >> 1], v << ((i & 1) * 16)); }

__device__ __forceinline__ uint32_t ldu(const uint32_t* __restrict__ S32, int x) {   // unaligned u32 at byte x
    const int w = x >> 2;
    return __funnelshift_r(__ldg(S32 + w), __ldg(S32 + w + 1), (x & 3) * 8);
}

__device__ __forceinline__ int ilog2x4(uint32_t v) {   // quarter-bit log2, 1 <= v < 2^30: 4*floor(log2 v) + next two mantissa bits
    const int n = 31 - __clz(v);
    return 4 * n + (int)(((v << 2) >> n) & 3u);
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

// number of equal leading bytes of two 16-byte strings given as four words each (16 = all equal)
__device__ __forceinline__ int eq16(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3) {
    const uint32_t w01 = x0 ? x0 : x1, w23 = x2 ? x2 : x3;
    const int b01 = x0 ? 0 : 4, b23 = x2 ? 8 : 12;
    const bool lo = (x0 | x1) != 0u;
    const uint32_t w = lo ? w01 : w23;
    const int b = lo ? b01 : b23;
    return b + (w ? (__ffs(w) - 1) >> 3 : 4);
}

// warp-cooperative: length of the common prefix of S[y..] and S[y-d..], up to maxn (128 bytes per step)
__device__ __forceinline__ int coop_match(const uint32_t* __restrict__ S32, int y, int d, int maxn, int lane) {
    int n = 0;
    while (n < maxn) {
        const int k = y + n + 4 * lane;
        const uint32_t x = ldu(S32, k) ^ ldu(S32, k - d);
        const uint32_t mism = __ballot_sync(kFull, x != 0);
        if (mism) {
            const int first = __ffs(mism) - 1;
            const uint32_t xx = __shfl_sync(kFull, x, first);
            n += 4 * first + ((__ffs(xx) - 1) >> 3);
            break;
        }
        n += 128;
    }
    return min(n, maxn);
}

// 16-bit mask of the bytes of a uint4 that differ from the replicated byte v4
__device__ __forceinline__ uint32_t ne_mask16(uint4 x, uint32_t v4) {
    const uint32_t a = (__vcmpne4(x.x, v4) & 0x08040201u) * 0x01010101u >> 24;
    const uint32_t b = (__vcmpne4(x.y, v4) & 0x08040201u) * 0x01010101u >> 24;
    const uint32_t c = (__vcmpne4(x.z, v4) & 0x08040201u) * 0x01010101u >> 24;
    const uint32_t d = (__vcmpne4(x.w, v4) & 0x08040201u) * 0x01010101u >> 24;
    return a | (b << 4) | (c << 8) | (d << 12);
}

// warp-cooperative: first position >= from (and < e) whose byte differs from v; e if none.  1 KiB per round trip.
__device__ __forceinline__ int run_end(const uint8_t* __restrict__ S, int from, int e, uint32_t v, int lane) {
    const uint32_t v4 = v * 0x01010101u;
    int a = from & ~15;
    int skip = from - a;                                  // bytes of the first 16 that precede `from`
    while (a < e) {
        const uint4* g = reinterpret_cast<const uint4*>(S + a) + lane;
        const uint4 x = __ldg(g), y = __ldg(g + 32);
        uint32_t mx = ne_mask16(x, v4);
        if (lane == 0) mx &= ~((1u << skip) - 1u);
        const uint32_t my = ne_mask16(y, v4);
        const uint32_t bx = __ballot_sync(kFull, mx != 0);
        if (bx) {
            const int f = __ffs(bx) - 1;
            const uint32_t mm = __shfl_sync(kFull, mx, f);
            return min(e, a + 16 * f + __ffs(mm) - 1);
        }
        const uint32_t by = __ballot_sync(kFull, my != 0);
        if (by) {
            const int f = __ffs(by) - 1;
            const uint32_t mm = __shfl_sync(kFull, my, f);
            return min(e, a + 512 + 16 * f + __ffs(mm) - 1);
        }
        a += 1024; skip = 0;
    }
    return e;
}

}  // namespace

template <class Cfg>
__global__ void __launch_bounds__(kLzWarps * 32) k_lz(BatchD B) {
    constexpr int HB3 = Cfg::HB3, HB6 = Cfg::HB6;
    using WarpMem = vcp::WarpMem<Cfg>;
    extern __shared__ __align__(16) unsigned char lz_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpMem& M = reinterpret_cast<WarpMem*>(lz_smem)[warp];
  // persistent warps: sub-chunks are handed out in stream order from a work queue, so the grid never has a ragged last wave
  for (;;) {
    int item = 0;
    if (lane == 0) item = (int)atomicAdd(&B.counters[0], 1u);
    item = __shfl_sync(kFull, item, 0);
    if (item >= B.nitems) return;
    if (B.lz_order) item = (int)B.lz_order[item];                             // heaviest sub-chunks first: the queue drains into cheap ones
    const int sub_first = (int)B.item2sub[item];
    const int sub_count = (item + 1 < B.nitems ? (int)B.item2sub[item + 1] : B.nsub) - sub_first;
   for (int si = 0; si < sub_count; si++) {
    const int sub = sub_first + si;
    const BlockD& blk = B.blocks[B.sub2blk[sub]];
    const PageD& pg = B.pages[blk.page];
    const uint8_t* __restrict__ S = pg.filt;
    const uint32_t* __restrict__ S32 = reinterpret_cast<const uint32_t*>(S);     // page streams are 256-byte aligned
    const int F = (int)pg.filt_len;
    const int bpp = pg.c;
    const int rowlen = 1 + pg.w * pg.c;                                          // bytes of one filtered row
    const int s = (int)blk.start + (sub - blk.sub0) * kSubBytes;
    const int e = min(s + kSubBytes, (int)(blk.start + blk.len));
    const int base = s - kMaxDist;                                               // table entries are pos - base (u16), 0 = empty
    uint32_t* __restrict__ tok = B.tokens + ((S - B.filt_base) + s);

    if (si == 0) {
        // ---- first sub-chunk of the group: clear tables + histogram, then prime (below)
        uint4* z = reinterpret_cast<uint4*>(&M);
        const uint4 zero = make_uint4(0, 0, 0, 0);
        for (int i = lane; i < (int)(sizeof(WarpMem) / 16); i += 32) z[i] = zero;
    } else {
        // ---- next sub-chunk of the group: the tables already hold every position the parse of the previous sub-chunk
        //      inserted; move them to the new base (s advanced by 32 KiB; what falls out of the window becomes empty)
        uint4* z = reinterpret_cast<uint4*>(&M);
        constexpr int kTableVec = (int)((sizeof(M.t3) + sizeof(M.t6)) / 16);
        for (int i = lane; i < kTableVec; i += 32) {
            uint4 v = z[i];
            // per u16 half: h >= 0x8000 ? h - 0x8000 : 0   ==   (h & 0x7FFF) masked by the replicated top bit
            #define VCP_REBASE(x) ((x) & 0x7FFF7FFFu & ((((x) >> 15) & 0x00010001u) * 0xFFFFu))
            v.x = VCP_REBASE(v.x); v.y = VCP_REBASE(v.y); v.z = VCP_REBASE(v.z); v.w = VCP_REBASE(v.w);
            #undef VCP_REBASE
            z[i] = v;
        }
        const uint4 zero = make_uint4(0, 0, 0, 0);
        uint4* hz = reinterpret_cast<uint4*>(M.hist);
        for (int i = lane; i < (int)(sizeof(M.hist) / 16); i += 32) hz[i] = zero;
    }
    __syncwarp();

    // ---- prime with the kPrimeBytes in front of the sub-chunk.  The history streams through 512-byte register blocks (one uint4
    //      per lane, two blocks ahead in flight).  A block is ONE insert step: a lane hashes its own 16 positions (its 16 bytes and
    //      the first 3 of its neighbour's) and reads their buckets; after one barrier the highest position of the block wins each
    //      bucket (atomicMax).  The state to reach is: newest way = highest position of the last block that touches the bucket,
    //      older way = highest position of the block that touched it before that (tests/model/deflate_model.c prime_win walks the
    //      blocks forwards and shifts the bucket once per block).  The blocks are walked BACKWARDS here, nearest first: the first
    //      block that touches a bucket fills the newest way, the second the older way, and every later one finds the bucket full
    //      and issues nothing — a shared-memory atomic holds the load/store pipe 2 cycles per lane, and walking forwards spent
    //      them on entries that the next blocks overwrote.  A block that is one repeated byte (white paper after filtering: most
    //      of a text page) hashes every position to the same two buckets: one store per table.
    if (si == 0) {
        const int h0 = max(0, s - kPrimeBytes);                                  // multiple of 512 (s and kPrimeBytes are)
        const uint4* __restrict__ S128 = reinterpret_cast<const uint4*>(S);
        const uint4 none = make_uint4(0, 0, 0, 0);
        uint4 bc = none, bn = none;
        uint32_t after0 = __ldg(S32 + (s >> 2));                                 // first word behind the block in hand
        if (s - 512 >= h0) bc = __ldg(S128 + ((s - 512) >> 4) + lane);
        if (s - 1024 >= h0) bn = __ldg(S128 + ((s - 1024) >> 4) + lane);
        for (int b0 = s - 512; b0 >= h0; b0 -= 512) {
            const uint4 bf = b0 - 1024 >= h0 ? __ldg(S128 + ((b0 - 1024) >> 4) + lane) : none;
            const uint32_t c00 = __shfl_sync(kFull, bc.x, 0);
            if (__all_sync(kFull, bc.x == c00 && bc.y == c00 && bc.z == c00 && bc.w == c00) && c00 == __funnelshift_l(c00, c00, 8) &&
                after0 == c00 && b0 + 516 <= F) {
                if (lane == 0) {
                    const uint32_t h3 = ((c00 & 0xFFFFFFu) * 0x9E3779B1u) >> (32 - HB3);
                    const uint32_t h6 = (c00 * 0x9E3779B1u) >> (32 - HB6);
                    const uint32_t pos = (uint32_t)(b0 + 511 - base);
                    const uint32_t w3 = M.t3[h3], w6 = M.t6[h6];
                    if (!(w3 >> 16)) M.t3[h3] = pos << 16; else if (!(w3 & 0xFFFFu)) M.t3[h3] = w3 | pos;
                    if (!(w6 >> 16)) M.t6[h6] = pos << 16; else if (!(w6 & 0xFFFFu)) M.t6[h6] = w6 | pos;
                }
                __syncwarp();
                after0 = c00; bc = bn; bn = bf;
                continue;
            }
            // this lane's 16 bytes and the 4 behind them (the last lane's neighbour is lane 0 of the block behind)
            uint32_t wd[5] = {bc.x, bc.y, bc.z, bc.w, 0u};
            {
                const uint32_t nx = __shfl_down_sync(kFull, bc.x, 1);
                wd[4] = lane == 31 ? after0 : nx;
            }
            const int q0 = b0 + 16 * lane;
            uint32_t hh[16], bb[16];
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const uint32_t cur4 = __funnelshift_r(wd[k >> 2], wd[(k >> 2) + 1], (k & 3) * 8);
                hh[k] = ((cur4 & 0xFFFFFFu) * 0x9E3779B1u) >> (32 - HB3);
                bb[k] = M.t3[hh[k]];
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const uint32_t pos = (uint32_t)(q0 + k - base);
                if (q0 + k + 2 < F && !(bb[k] & 0xFFFFu))
                    atomicMax(&M.t3[hh[k]], (bb[k] >> 16) ? (bb[k] | pos) : (pos << 16));
            }
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const uint32_t cur4 = __funnelshift_r(wd[k >> 2], wd[(k >> 2) + 1], (k & 3) * 8);
                hh[k] = (cur4 * 0x9E3779B1u) >> (32 - HB6);
                bb[k] = M.t6[hh[k]];
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const uint32_t pos = (uint32_t)(q0 + k - base);
                if (q0 + k + kH2Bytes <= F && !(bb[k] & 0xFFFFu))
                    atomicMax(&M.t6[hh[k]], (bb[k] >> 16) ? (bb[k] | pos) : (pos << 16));
            }
            __syncwarp();
            after0 = c00; bc = bn; bn = bf;
        }
    }

    // ---- main loop.  Register window: three 128-byte chunks at A0, A0+128, A0+256 (A0 multiple of 128).
    int p = s;
    uint32_t ntok = 0;
    int score = 0;                                                               // EMA of literal tokens per window, x8
    uint32_t cost_ep = 0; int lgN = 0;                                           // cost epoch of M.clg, and log2 of the token count it was taken at
    int A0 = ((p - 4) >> 7) << 7;                                                // may be -128: the pad in front of the stream is addressable
    uint32_t w0r = __ldg(S32 + (A0 >> 2) + lane), w1r = __ldg(S32 + (A0 >> 2) + 32 + lane), w2r = __ldg(S32 + (A0 >> 2) + 64 + lane);
    while (p < e) {
        if ((ntok >> kCostEpochLog2) != cost_ep) {
            // short matches are priced against literals with the token statistics of this sub-chunk: log2 tables, refreshed every 64
            // tokens (exact counts per window cost 8 % of the kernel's instructions for 0.1 % of PNG size)
            cost_ep = ntok >> kCostEpochLog2; lgN = ilog2x4(ntok + 1);
            for (int i = lane; i < kHistSize; i += 32) M.clg[i] = (uint8_t)ilog2x4(hist_get(M, i) + 1);
            __syncwarp();
        }
        int ofs = p - 4 - A0;
        if (ofs >= 384) {                                                        // long jump: refill
            A0 = ((p - 4) >> 7) << 7; ofs = p - 4 - A0;
            w0r = __ldg(S32 + (A0 >> 2) + lane); w1r = __ldg(S32 + (A0 >> 2) + 32 + lane); w2r = __ldg(S32 + (A0 >> 2) + 64 + lane);
        } else {
            while (ofs >= 128) {                                                 // slide: the prefetched chunk becomes "next"
                w0r = w1r; w1r = w2r; A0 += 128; ofs -= 128;
                w2r = (A0 + 256 < F + 128) ? __ldg(S32 + (A0 >> 2) + 64 + lane) : 0u;
            }
        }
        const int q = p + lane;
        // V[l] = word (ofs>>2) + l of the register window: bytes [p-4-r, p-4-r+128), r = ofs & 3
        uint32_t V;
        {
            const int kb = (ofs >> 2) + lane;
            const uint32_t x = __shfl_sync(kFull, w0r, kb & 31), y = __shfl_sync(kFull, w1r, kb & 31);
            V = (kb & 32) ? y : x;
        }
        const int o = (ofs & 3) + lane, k = o >> 2, sh = (o & 3) * 8;
        const uint32_t a0 = __shfl_sync(kFull, V, k), a1 = __shfl_sync(kFull, V, k + 1), a2 = __shfl_sync(kFull, V, k + 2);
        const uint32_t a3 = __shfl_sync(kFull, V, k + 3), a4 = __shfl_sync(kFull, V, k + 4), a5 = __shfl_sync(kFull, V, k + 5);
        const uint32_t lo = __funnelshift_r(a0, a1, sh), cur4 = __funnelshift_r(a1, a2, sh), nxt4 = __funnelshift_r(a2, a3, sh);
        const uint32_t q8 = __funnelshift_r(a3, a4, sh), q12 = __funnelshift_r(a4, a5, sh);   // bytes [q+8, q+16)
        const uint32_t c0 = __shfl_sync(kFull, V, k + 8), c1 = __shfl_sync(kFull, V, k + 9);
        const uint32_t lo2 = __funnelshift_r(c0, c1, sh);                        // bytes [q+28, q+32)
        const uint32_t b2 = (c1 >> sh) & 0xFFu;                                  // byte q+32
        // equality bits for distance 1 and distance bpp at positions q and q+32
        const uint32_t bq = cur4 & 0xFFu;
        const bool e1a = (bq == (lo >> 24)) && q >= 1;
        const bool eba = (bq == ((lo >> (8 * (4 - bpp))) & 0xFFu)) && q >= bpp;
        const bool e1b = (b2 == (lo2 >> 24));
        const bool ebb = (b2 == ((lo2 >> (8 * (4 - bpp))) & 0xFFu));
        const unsigned long long m1 = (unsigned long long)__ballot_sync(kFull, e1a) | ((unsigned long long)__ballot_sync(kFull, e1b) << 32);
        const unsigned long long mb = (unsigned long long)__ballot_sync(kFull, eba) | ((unsigned long long)__ballot_sync(kFull, ebb) << 32);

        const int limit = min(kMaxMatch, e - q);                                 // <= 0 for lanes past the sub-chunk
        int bl = 0, bd = 0; bool bcap = false;
        // A lane compares its candidates only 16 bytes deep.  Fully compared ones are ranked in xbest (longer, then nearer);
        // the ones that hit the cap are remembered in capm and only measured to the end — cooperatively, 128 bytes per
        // round trip — if the greedy parse actually starts a token at this lane.
        uint32_t xbest = 0;                                                       // (length << 16) | (32768 - distance)
        uint32_t cbest = 0;                                                       // nearest capped candidate: ((32768 - distance) << 8) | length so far
        uint32_t capm = 0;                                                        // bits 0..4: table ways / row probe, 5: distance-1 run, 6: distance-bpp run
        int cpos[5];
#pragma unroll
        for (int w = 0; w < 5; w++) cpos[w] = q;
        const uint32_t h3 = ((cur4 & 0xFFFFFFu) * 0x9E3779B1u) >> (32 - HB3);
        const uint32_t h6 = (cur4 * 0x9E3779B1u) >> (32 - HB6);
        const uint32_t b3 = M.t3[h3], b6 = M.t6[h6];
        if (limit >= 3) {
            const int runcap = min(64 - lane, limit);
            {   // distance 1
                const unsigned long long inv = ~(m1 >> lane);
                const int r = inv ? __ffsll((long long)inv) - 1 : 64;
                const int l = min(r, runcap);
                const bool c = (l == runcap) && (runcap < limit);
                if (l >= 3) { if (c) { capm |= 32u; cbest = (32767u << 8) | (uint32_t)l; } else xbest = ((uint32_t)l << 16) | 32767u; }
            }
            if (bpp > 1) {
                const unsigned long long inv = ~(mb >> lane);
                const int r = inv ? __ffsll((long long)inv) - 1 : 64;
                const int l = min(r, runcap);
                const bool c = (l == runcap) && (runcap < limit);
                if (l >= 3) {
                    if (c) { capm |= 64u; cbest = max(cbest, ((uint32_t)(32768 - bpp) << 8) | (uint32_t)l); }
                    else xbest = max(xbest, ((uint32_t)l << 16) | (uint32_t)(32768 - bpp));
                }
            }
            if (!capm) {                                                          // a capped run outranks every hash candidate
                const int hcap = min(kLaneCap, limit);
                const bool ok6 = q + kH2Bytes <= F;
                const bool noisy = score >= Cfg::NOISY;                               // literal-dense stretch: older ways rarely pay for their loads
                int clen[5]; bool live[5];
#pragma unroll
                for (int w = 0; w < 5; w++) {
                    // w = 0..3: the two ways of the two tables; w = 4: the byte one filtered row up (smooth shading repeats there)
                    const uint32_t cnd = (w == 0) ? (b3 >> 16) : (w == 1) ? (b3 & 0xFFFFu) : (w == 2) ? (b6 >> 16) : (b6 & 0xFFFFu);
                    const int cp = w < 4 ? base + (int)cnd : q - rowlen;
                    const int d = q - cp;
                    live[w] = w < 4 ? (cnd != 0 && (w < 2 || ok6) && d > 0 && d <= kMaxDist && !(noisy && (w & 1)))
                                    : (Cfg::PROBE != 0 && rowlen <= kMaxDist && cp >= 0 && (Cfg::PROBE == 2 || !e1a));   // default: only where a run starts (inside a run distance 1 already serves)
                    cpos[w] = live[w] ? cp : q;
                    clen[w] = 0;
                }
                // stage A: the first 4 bytes only (two words per candidate, all ten loads issued before the first use) —
                // in noisy rows almost every candidate dies here
                uint32_t g0k[5], g1k[5];
#pragma unroll
                for (int w = 0; w < 5; w++) { const uint32_t* g = S32 + (cpos[w] >> 2); g0k[w] = __ldg(g); g1k[w] = __ldg(g + 1); }
                bool any = false;
#pragma unroll
                for (int w = 0; w < 5; w++) {
                    const uint32_t x0 = cur4 ^ __funnelshift_r(g0k[w], g1k[w], (cpos[w] & 3) * 8);
                    const int n = x0 ? (__ffs(x0) - 1) >> 3 : 4;
                    clen[w] = live[w] ? min(n, hcap) : 0;
                    live[w] = live[w] && x0 == 0u && hcap > 4;
                    any |= live[w];
                }
                // stage B: bytes 4..15 of the survivors (loads first, then the compares)
                if (any) {
                    uint32_t g2k[5], g3k[5], g4k[5];
#pragma unroll
                    for (int w = 0; w < 5; w++) {
                        g2k[w] = g3k[w] = g4k[w] = 0u;
                        if (live[w]) { const uint32_t* g = S32 + (cpos[w] >> 2); g2k[w] = __ldg(g + 2); g3k[w] = __ldg(g + 3); g4k[w] = __ldg(g + 4); }
                    }
#pragma unroll
                    for (int w = 0; w < 5; w++) {
                        if (live[w]) {
                            const int shc = (cpos[w] & 3) * 8;
                            const int n = eq16(0u, nxt4 ^ __funnelshift_r(g1k[w], g2k[w], shc), q8 ^ __funnelshift_r(g2k[w], g3k[w], shc),
                                               q12 ^ __funnelshift_r(g3k[w], g4k[w], shc));
                            clen[w] = min(n, hcap);
                        }
                    }
                }
#pragma unroll
                for (int w = 0; w < 5; w++) {
                    const int l = clen[w], d = q - cpos[w];
                    if (l >= 3) {
                        if (l == hcap && hcap < limit) { capm |= 1u << w; cbest = max(cbest, ((uint32_t)(32768 - d) << 8) | (uint32_t)l); }
                        else xbest = max(xbest, ((uint32_t)l << 16) | (uint32_t)(32768 - d));
                    }
                }
            }
            bcap = capm != 0u;
            if (bcap) { bl = (int)(cbest & 0xFFu); bd = 32768 - (int)(cbest >> 8); }
            else if (xbest) { bl = (int)(xbest >> 16); bd = 32768 - (int)(xbest & 0xFFFFu); }
            // price short matches against literals with the histogram as of the window start
            if (bl >= 3 && bl <= kCostMaxLen && ntok >= kCostWarm) {
                int lit = 0;
#pragma unroll
                for (int kk = 0; kk < kCostMaxLen; kk++) {
                    if (kk < bl) {
                        const uint32_t byte = (kk < 4 ? (cur4 >> (8 * kk)) : (nxt4 >> (8 * (kk - 4)))) & 0xFFu;
                        lit += lgN - (int)M.clg[byte];
                    }
                }
                const int ls = bl - 3, ds = dist_sym(bd);                         // bl <= 8: no length extra bits
                const int dx = ds < 4 ? 0 : (ds >> 1) - 1;
                const int mc = (lgN - (int)M.clg[257 + ls]) + (lgN - (int)M.clg[286 + ds]) + 4 * dx;
                if (mc >= lit) { bl = 0; bd = 0; }
            }
        }
        // ---- one-step lazy rule between neighbouring lanes
        {
            const int nxt = __shfl_down_sync(kFull, bl, 1);
            if (lane < 31 && bl >= 3 && bl < Cfg::LAZY && nxt > bl) { bl = 0; bd = 0; }
        }
        // ---- greedy parse from lane 0 by pointer jumping; whenever the parse starts a token on a lane with capped candidates,
        //      those are measured to the end (the best fully compared one competes too) and the parse is redone from there
        const int nvalid = min(32, e - p);
        uint32_t sel;
        for (;;) {
            int J = lane + (bl ? bl : 1);
            if (J >= nvalid) J = 32;
            uint32_t Msel = 1u << lane;
#pragma unroll
            for (int r = 0; r < 5; r++) {
                const uint32_t Mj = __shfl_sync(kFull, Msel, J & 31);
                const int Jj = __shfl_sync(kFull, J, J & 31);
                if (J < 32) { Msel |= Mj; J = Jj; }
            }
            sel = __shfl_sync(kFull, Msel, 0);
            const uint32_t cm = __ballot_sync(kFull, bcap) & sel;
            if (!cm) break;
            const int f = __ffs(cm) - 1;
            const int qf = p + f;
            const int limf = min(kMaxMatch, e - qf);
            const int rcf = min(64 - f, limf), hcf = min(kLaneCap, limf);
            const uint32_t capf = __shfl_sync(kFull, capm, f);
            uint32_t bestk = __shfl_sync(kFull, xbest, f);
            if (capf & 0x1Fu) {
                // table / row candidates: the next 128 bytes of all of them in ONE round trip (the q side is shared)
                const int off = hcf + 4 * lane;
                const uint32_t xq = ldu(S32, qf + off);
                int dw[5]; uint32_t xw[5];
#pragma unroll
                for (int w = 0; w < 5; w++) {
                    const int cpw = __shfl_sync(kFull, cpos[w], f);
                    dw[w] = qf - cpw;
                    xw[w] = ((capf >> w) & 1u) ? (xq ^ ldu(S32, cpw + off)) : 0u;
                }
#pragma unroll
                for (int w = 0; w < 5; w++) {
                    if ((capf >> w) & 1u) {
                        const uint32_t mism = __ballot_sync(kFull, xw[w] != 0u);
                        int Lw = hcf + 128;
                        if (mism) {
                            const int first = __ffs(mism) - 1;
                            const uint32_t xx = __shfl_sync(kFull, xw[w], first);
                            Lw = hcf + 4 * first + ((__ffs(xx) - 1) >> 3);
                        } else if (Lw < limf) {
                            Lw += coop_match(S32, qf + Lw, dw[w], limf - Lw, lane);      // rare: more than 144 equal bytes
                        }
                        bestk = max(bestk, ((uint32_t)min(Lw, limf) << 16) | (uint32_t)(32768 - dw[w]));
                    }
                }
            }
#pragma unroll
            for (int w = 5; w < 7; w++) {                                         // capped runs (then no table candidate was looked at)
                if ((capf >> w) & 1u) {
                    const int d = (w == 5) ? 1 : bpp;
                    const int Lw = rcf + coop_match(S32, qf + rcf, d, limf - rcf, lane);
                    bestk = max(bestk, ((uint32_t)Lw << 16) | (uint32_t)(32768 - d));
                }
            }
            if (lane == f) { bl = (int)(bestk >> 16); bd = 32768 - (int)(bestk & 0xFFFFu); bcap = false; }
        }
        const int last = 31 - __clz(sel);
        const int Ll = __shfl_sync(kFull, bl, last);
        const int dl = __shfl_sync(kFull, bd, last);
        const int ql = p + last;
        __syncwarp();                                                            // all histogram reads of this window are done
        // ---- emit tokens + histogram
        if ((sel >> lane) & 1u) {
            const int rank = __popc(sel & ((1u << lane) - 1u));
            if (bl) {
                tok[ntok + rank] = 0x80000000u | ((uint32_t)(bd - 1) << 8) | (uint32_t)(bl - 3);
                hist_add(M, 257 + len_sym(bl), 1u);
                hist_add(M, 286 + dist_sym(bd), 1u);
            } else {
                tok[ntok + rank] = bq;
                hist_add(M, (int)bq, 1u);
            }
        }
        ntok += __popc(sel);
        score = score - (score >> 3) + __popc(__ballot_sync(kFull, ((sel >> lane) & 1u) && bl == 0));
        int next = ql + (Ll ? Ll : 1);
        // ---- a maximal distance-1 run goes on: measure it 1 KiB per round trip, emit the full 258-tokens it holds
        if (Ll == kMaxMatch && dl == 1 && e - next >= kMaxMatch) {
            const uint32_t v = __ldg(S + next - 1);
            const int rend = run_end(S, next, e, v, lane);
            const uint32_t extra = (uint32_t)((rend - next) / kMaxMatch);
            for (uint32_t i = lane; i < extra; i += 32) tok[ntok + i] = 0x80000000u | (uint32_t)(kMaxMatch - 3);
            if (extra) {
                if (lane == 0) { hist_add(M, 257 + 28, extra); hist_add(M, 286, extra); }
                ntok += extra;
                next += (int)extra * kMaxMatch;
            }
        }
        // ---- insert this window's positions (atomicMax: the highest position of a bucket group wins).
        //      If bytes p .. p+34 are one repeated byte every lane has the same two buckets: only the highest valid lane writes.
        {
            const uint32_t pos = (uint32_t)(q - base);
            const bool i3 = q < e && q + 2 < F, i6 = q < e && q + kH2Bytes <= F;
            const bool uniform = ((m1 >> 1) & 0x3FFFFFFFFull) == 0x3FFFFFFFFull;
            if (uniform) {
                const uint32_t m3 = __ballot_sync(kFull, i3), m6 = __ballot_sync(kFull, i6);
                if (i3 && (m3 >> lane) == 1u) M.t3[h3] = (pos << 16) | (b3 >> 16);
                if (i6 && (m6 >> lane) == 1u) M.t6[h6] = (pos << 16) | (b6 >> 16);
            } else {
                if (i3) atomicMax(&M.t3[h3], (pos << 16) | (b3 >> 16));
                if (i6) atomicMax(&M.t6[h6], (pos << 16) | (b6 >> 16));
            }
        }
        __syncwarp();
        p = next;
    }
    // ---- results
    if (lane == 0) B.sub_ntok[sub] = ntok;
    uint32_t* hout = B.sub_hist + (size_t)sub * kHistSize;
    for (int i = lane; i < kHistSize; i += 32) hout[i] = hist_get(M, i);
    __syncwarp();
   }
  }
}

// Work items sorted by a cost hint, heaviest first (longest-processing-time-first scheduling): a sub-chunk costs roughly in
// proportion to the ink of the rows it covers (the PNG filter's winning |residual| sum; blank rows are one long run).
// Stable counting sort with 64 keys: k_lz_keys bins the items (global histogram in B.counters[64..128), one histogram per CTA
// behind the keys), k_lz_scatter places them.  The order only changes scheduling, never the output.
constexpr int kOrderKeys = 64;
#ifndef VCP_ORDER_GROUP
#define VCP_ORDER_GROUP 8
#endif
constexpr int kOrderGroup = VCP_ORDER_GROUP;              // consecutive sub-chunks that share one key (the heaviest of the group): a sub-chunk primes
                                                          // its tables from the 16 KiB in front of it, which its neighbour reads at the same moment

__device__ __forceinline__ int lz_item_key(const BatchD& B, int item) {
    const int sub = (int)B.item2sub[item];
    const BlockD& blk = B.blocks[B.sub2blk[sub]];
    const PageD& pg = B.pages[blk.page];
    const int rowlen = 1 + pg.w * pg.c;
    const int s = (int)blk.start + (sub - blk.sub0) * kSubBytes;
    const int e = min(s + kSubBytes, (int)(blk.start + blk.len));
    const int y0 = s / rowlen, y1 = (e - 1) / rowlen;
    const int n = y1 - y0 + 1, step = max(1, n / 8);
    int busy = 0, seen = 0;
    for (int y = y0; y <= y1 && seen < 8; y += step, seen++) busy += B.row_busy[pg.row0 + y];
    return seen ? min(kOrderKeys - 1, (busy * 8 / seen + 31) >> 5) : 0;       // scaled to 8 rows
}

__global__ void __launch_bounds__(256) k_lz_keys(BatchD B, uint8_t* __restrict__ keys, uint32_t* __restrict__ blkcnt) {
    __shared__ uint32_t cnt[kOrderKeys];
    if (threadIdx.x < kOrderKeys) cnt[threadIdx.x] = 0;
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int k = i < B.nitems ? lz_item_key(B, i) : 0;
    for (int d = 1; d < kOrderGroup; d <<= 1) k = max(k, __shfl_xor_sync(0xffffffffu, k, d));     // neighbours travel together
    if (i < B.nitems) { keys[i] = (uint8_t)k; atomicAdd(&cnt[k], 1u); }
    __syncthreads();
    if (threadIdx.x < kOrderKeys) {
        blkcnt[(size_t)blockIdx.x * kOrderKeys + threadIdx.x] = cnt[threadIdx.x];
        if (cnt[threadIdx.x]) atomicAdd(&B.counters[64 + threadIdx.x], cnt[threadIdx.x]);
    }
}

// Stable placement: inside a key the items keep stream order, so the sub-chunks a warp primes its tables from (the 16 KiB in front of
// its own) were parsed moments before by a neighbouring warp and are still in L2, and so are the rows above for the row probe.
__global__ void __launch_bounds__(256) k_lz_scatter(BatchD B, const uint8_t* __restrict__ keys, const uint32_t* __restrict__ blkcnt) {
    __shared__ uint32_t base[kOrderKeys];                   // first slot of this CTA's items of every key (heaviest key first)
    __shared__ uint32_t wcnt[8][kOrderKeys];                // per warp: items of the key in the warps in front of it
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = threadIdx.x; j < 8 * kOrderKeys; j += 256) (&wcnt[0][0])[j] = 0;
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = i < B.nitems ? (int)keys[i] : 255;
    const uint32_t same = __match_any_sync(0xffffffffu, k);
    const int lower = __popc(same & ((1u << lane) - 1u));
    if (k != 255 && lower == 0) wcnt[warp][k] = (uint32_t)__popc(same);
    __syncthreads();
    if (threadIdx.x < kOrderKeys) {
        const int kk = threadIdx.x;
        uint32_t o = 0;
        for (int q = kOrderKeys - 1; q > kk; q--) o += B.counters[64 + q];          // heavier keys come first
        for (int b2 = 0; b2 < (int)blockIdx.x; b2++) o += blkcnt[(size_t)b2 * kOrderKeys + kk];
        base[kk] = o;
        uint32_t run = 0;
        for (int w = 0; w < 8; w++) { const uint32_t t = wcnt[w][kk]; wcnt[w][kk] = run; run += t; }
    }
    __syncthreads();
    if (k != 255) B.lz_order[base[k] + wcnt[warp][k] + lower] = (uint32_t)i;
}

int launch_lz_order(const BatchD& b, cudaStream_t st) {
    if (b.nitems == 0 || !b.lz_order || !b.row_busy) return 0;
    uint8_t* keys = reinterpret_cast<uint8_t*>(b.lz_order + b.nitems);         // the order buffer has room for the keys and the per-CTA counts behind it
    uint32_t* blkcnt = reinterpret_cast<uint32_t*>(keys + ((b.nitems + 15) & ~15));
    const int ctas = (b.nitems + 255) / 256;
    k_lz_keys<<<ctas, 256, 0, st>>>(b, keys, blkcnt);
    k_lz_scatter<<<ctas, 256, 0, st>>>(b, keys, blkcnt);
    return 2;
}

template <class Cfg>
static int launch_lz_cfg(const BatchD& b, cudaStream_t st) {
    const size_t smem = sizeof(WarpMem<Cfg>) * kLzWarps;
    // CTAs the device can hold at once (persistent grid).  The dynamic shared-memory limit is a per-device function attribute:
    // one process may drive several GPUs (prepare_pages_all_gpus), so the cache is keyed by device ordinal.
    static std::atomic<int> resident_by_dev[kMaxDevices];
    int dev = 0;
    cudaGetDevice(&dev);
    const int slot = std::min(std::max(dev, 0), kMaxDevices - 1);
    int resident = resident_by_dev[slot].load(std::memory_order_acquire);
    if (!resident || dev >= kMaxDevices) {
        cudaFuncSetAttribute(k_lz<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int per_sm = 0, sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_lz<Cfg>, kLzWarps * 32, smem);
        resident = std::max(1, per_sm) * sms;
        resident_by_dev[slot].store(resident, std::memory_order_release);
    }
    const int ctas = std::min((b.nitems + kLzWarps - 1) / kLzWarps, resident);
    k_lz<Cfg><<<ctas, kLzWarps * 32, smem, st>>>(b);
    return 1;
}

int launch_lz(const BatchD& b, cudaStream_t st) {
    if (b.nitems == 0) return 0;
    if (b.level >= 7) return launch_lz_cfg<CfgBest>(b, st);
    if (b.level >= 1 && b.level <= 3) return launch_lz_cfg<CfgFast>(b, st);
    return launch_lz_cfg<CfgDefault>(b, st);                  // 4..6 (and 0: stored blocks are decided later, tokens unused)
}

}  // namespace vcp
