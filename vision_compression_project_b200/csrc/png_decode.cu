// png_decode.cu — zlib inflate and PNG un-filtering on the GPU (the reverse of deflate_*.cu / png_filter.cu).
//
// Replaces what Pillow runs for Image.open(png).load(): zlib inflate() driven by libImaging/ZipDecode.c and its per-row
// un-filter (None/Sub/Up/Avg/Paeth).  SURVEY.md §8 f-4: re-reading the reference's images/page_###.png cache
// (backend/README.md:238-243) and validating this library's own PNGs at speed.  Restated for tests by
// oracle/restate.py:png_unfilter and Python's zlib.
//
// Inflate.  A deflate stream is serial, but a PNG cuts it into IDAT chunks, and an encoder that ends every chunk on a block
// boundary (this library: one byte-aligned deflate block per IDAT) leaves chunks that can be parsed independently.  Whether a
// foreign PNG has that property is *checked*, not assumed:
//   k_infl_probe   one warp per IDAT: parse it as if it began a block, count the bytes it produces, and record whether it ended
//                  exactly on its last byte at a block boundary.  IDAT 0 really does begin a block, so if every IDAT passes, by
//                  induction every IDAT begins one (no speculation left) and the page is inflated segment-parallel.
//   k_infl_plan    one warp per page: all IDATs passed and the byte counts add up -> output offsets per IDAT, mode = parallel.
//   k_infl_seg     one warp per IDAT of a parallel page.  Matches may reach up to 32 KiB behind the IDAT's first byte, into output
//                  another warp is still producing, so the warp inflates *symbolically*: 16-bit elements, 0..255 = a byte,
//                  256 + j = "byte j of the 32 KiB in front of this IDAT".  Copies move symbols like bytes.  The 32 Ki-entry window
//                  is a ring in shared memory (64 KiB per warp), so a match costs a shared-memory round trip, not a global one.
//   k_infl_window  one CTA per page walks the IDATs in order and makes the last 32 KiB of each concrete (each step a 32 Ki-wide
//                  gather through the previous, already concrete, window).
//   k_infl_resolve every other position of every IDAT in parallel: symbol -> byte through the window in front of its IDAT.
//   k_inflate      the serial path for everything else (Pillow's 64 KiB IDAT cuts fall mid-block): one warp per page, lane 0 parses
//                  32 tokens at a time, the warp executes them against a 32 KiB byte ring in shared memory.
// Un-filter.  Pixel (x, y) needs (x-1, y), (x, y-1), (x-1, y-1): a wavefront.
//   k_unfilter     one warp per band of 32 rows, lane l on row y0 + l running l pixels behind lane l - 1, so the pixel above
//                  arrives by one shuffle per step; rows are staged through shared memory 32 pixels at a time with coalesced
//                  loads and stores.  Bands of a page run concurrently two chunks apart: lane 0 reads the last row of the band
//                  above from global memory once that band's progress flag (release / acquire) covers it.  CTAs take their band
//                  range from a ticket, so a band only ever waits on bands that are already running.
// Every loop is bounded by the stream / output length: malformed input ends in a status, never in a hang.
#include "vcp_internal.cuh"

namespace vcp {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kFastBits = 10;
constexpr int kWin = 32768;
constexpr int kZWords = 512;              // compressed-input ring (words) the warp keeps filled ahead of lane 0's parser
constexpr int kResolveChunk = 32768;      // positions per CTA of k_infl_resolve (keep in step with api.cu)

enum { INF_OK = 0, INF_BAD_HEADER = -101, INF_BAD_BLOCK = -102, INF_BAD_CODE = -103, INF_OVERRUN = -104, INF_BAD_DIST = -105,
       INF_SHORT = -106, INF_BAD_FILTER = -107, INF_SEG = -108 };

struct InflMem {
    uint16_t fast_ll[1 << kFastBits];     // (symbol << 4) | length, 0 = not in the fast table
    uint16_t fast_d[1 << kFastBits];
    uint16_t sorted_ll[288], sorted_d[32];
    uint16_t cnt_ll[16], cnt_d[16];
    uint8_t lens[320];
    uint32_t tok[32];                     // one batch of tokens: literal = bit 31 | byte, match = distance << 9 | length
    uint32_t zbuf[kZWords];               // the compressed stream in front of the parser: word w of the stream at zbuf[w % kZWords]
};

// LSB-first bit reader (lane 0) over a byte stream that has >= 64 addressable bytes of slack behind it.  It refills 32 bits at a
// time from two aligned words: one already in a register, the other from the shared-memory ring the whole warp keeps filled
// (a global load here would put a DRAM round trip in front of every second token).
struct BitReader {
    const uint32_t* zw; int mis;                 // aligned word pointer of the stream start, byte misalignment 0..3
    const uint32_t* zs;                          // InflMem::zbuf
    unsigned long long n, pos;                   // stream length, bytes fetched so far
    unsigned long long buf; int cnt; bool over;
    uint32_t nextw;                              // aligned word (pos + mis) / 4
    uint32_t maxw;                               // last word index that may be read
    __device__ void init(const uint8_t* z, unsigned long long len, const uint32_t* ring) {
        mis = (int)((uintptr_t)z & 3); zw = reinterpret_cast<const uint32_t*>(z - mis); zs = ring;
        n = len; pos = 0; buf = 0; cnt = 0; over = false;
        maxw = (uint32_t)((len + mis + 60) >> 2);
        nextw = __ldg(zw);
    }
    __device__ uint32_t word() const { return (uint32_t)((pos + mis) >> 2); }
    __device__ void refill() {
        if (cnt <= 32) {
            const uint32_t w = word();
            const uint32_t hi = zs[(w + 1) & (kZWords - 1)];
            const uint32_t v = __funnelshift_r(nextw, hi, 8 * (int)((pos + mis) & 3));   // bytes [pos, pos + 4)
            nextw = hi;
            buf |= (unsigned long long)v << cnt; cnt += 32; pos += 4;
            if (pos > n + 12) over = true;
        }
    }
    __device__ uint32_t peek(int k) { return (uint32_t)(buf & ((1ull << k) - 1ull)); }
    __device__ void drop(int k) { buf >>= k; cnt -= k; }
    __device__ uint32_t bits(int k) { if (cnt < k) refill(); const uint32_t v = peek(k); drop(k); return v; }
    __device__ unsigned long long bits_used() const { return pos * 8ull - (unsigned long long)cnt; }
    __device__ unsigned long long byte_pos() const { return pos - (unsigned long long)(cnt >> 3); }   // first byte not consumed (cnt % 8 == 0)
    __device__ void seek(unsigned long long p) { pos = p; buf = 0; cnt = 0; nextw = __ldg(zw + ((p + mis) >> 2)); }
};

// Whole warp: load stream words [filled, ...) into the ring until it reaches kZWords ahead of word w0 (lane 0's position).
__device__ __forceinline__ void topup(InflMem& M, const BitReader& br, uint32_t& filled, uint32_t w0) {
    const int lane = threadIdx.x & 31;
    if (filled < w0) filled = w0;
    while (filled + 32 <= w0 + kZWords) {
        const uint32_t i = filled + lane;
        M.zbuf[i & (kZWords - 1)] = __ldg(br.zw + min(i, br.maxw));
        filled += 32;
    }
    __syncwarp();
}

__constant__ uint16_t kLenBase[29] = {3,4,5,6,7,8,9,10,11,13,15,17,19,23,27,31,35,43,51,59,67,83,99,115,131,163,195,227,258};
__constant__ uint8_t kLenExtra[29] = {0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0};
__constant__ uint16_t kDistBase[30] = {1,2,3,4,5,7,9,13,17,25,33,49,65,97,129,193,257,385,513,769,1025,1537,2049,3073,4097,6145,8193,12289,16385,24577};
__constant__ uint8_t kDistExtra[30] = {0,0,0,0,1,1,2,2,3,3,4,4,5,5,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13};
__constant__ uint8_t kClOrd[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// Build canonical decode structures for n symbols with code lengths L (one thread).  Returns false for an over-subscribed code.
__device__ bool build_table(const uint8_t* L, int n, uint16_t* cnt, uint16_t* sorted, uint16_t* fast) {
    for (int i = 0; i < 16; i++) cnt[i] = 0;
    for (int i = 0; i < n; i++) cnt[L[i]]++;
    cnt[0] = 0;
    int left = 1;
    for (int l = 1; l < 16; l++) { left <<= 1; left -= cnt[l]; if (left < 0) return false; }
    uint16_t offs[16]; offs[1] = 0;
    for (int l = 1; l < 15; l++) offs[l + 1] = offs[l] + cnt[l];
    for (int i = 0; i < n; i++) if (L[i]) sorted[offs[L[i]]++] = (uint16_t)i;
    for (int i = 0; i < (1 << kFastBits); i++) fast[i] = 0;
    // fast table: canonical codes of length <= kFastBits, bit-reversed (the stream is LSB first)
    uint32_t code = 0; int idx = 0;
    for (int l = 1; l <= kFastBits; l++) {
        for (int k = 0; k < cnt[l]; k++, idx++, code++) {
            const uint32_t r = __brev(code) >> (32 - l);
            for (uint32_t e = r; e < (1u << kFastBits); e += 1u << l) fast[e] = (uint16_t)((sorted[idx] << 4) | l);
        }
        code <<= 1;
    }
    return true;
}

// one symbol (lane 0 only)
__device__ int decode_sym(BitReader& br, const uint16_t* fast, const uint16_t* cnt, const uint16_t* sorted) {
    if (br.cnt < 15) br.refill();
    const uint16_t e = fast[br.peek(kFastBits)];
    if (e) { br.drop(e & 15); return e >> 4; }
    int code = 0, first = 0, index = 0;
    unsigned long long b = br.buf;
    for (int l = 1; l <= 15; l++) {
        code |= (int)(b & 1ull); b >>= 1;
        const int c = cnt[l];
        if (code - c < first) { br.drop(l); return sorted[index + (code - first)]; }
        index += c; first += c; first <<= 1; code <<= 1;
    }
    return -1;
}


// What a warp does with the tokens it parses.
//   COUNT: nothing but add up their sizes (k_infl_probe).
//   BYTES: execute against a byte ring (k_inflate).
//   SYMS:  execute against a ring of 16-bit symbols; sources in front of the stream become 256 + window index (k_infl_seg).
enum { COUNT = 0, BYTES = 1, SYMS = 2 };

// Inflates deflate blocks from `br` until the final block (until_final) or until a block ends exactly on the last bit of the
// reader's range (!until_final).  `abs0` is the position of out[0] in the page's filtered stream (distance check).  All lanes
// call it; lane 0 parses.  Returns the status; *produced = bytes written (all lanes), *final_seen = the BFINAL block was met.
template <int MODE, typename T>
__device__ int inflate_blocks(InflMem& M, T* __restrict__ ring, BitReader& br, const uint8_t* __restrict__ zbase, T* __restrict__ out,
                              unsigned long long cap, unsigned long long abs0, bool until_final, unsigned long long* produced, bool* final_seen) {
    const int lane = threadIdx.x & 31;
    constexpr uint32_t RM = kWin - 1;
    unsigned long long pos = 0;
    int status = INF_OK;
    int last = 0;
    bool done = false;
    uint32_t filled = 0;                      // stream words [.., filled) have been put into M.zbuf (warp-uniform)
    while (status == INF_OK && !done) {
        topup(M, br, filled, __shfl_sync(kFull, br.word(), 0));       // a block header is at most ~330 bytes
        int btype = 0;
        if (lane == 0) { last = (int)br.bits(1); btype = (int)br.bits(2); if (br.over) status = INF_SHORT; }
        status = __shfl_sync(kFull, status, 0);
        if (status != INF_OK) break;
        last = __shfl_sync(kFull, last, 0); btype = __shfl_sync(kFull, btype, 0);
        if (btype == 0) {
            // stored: byte-align, LEN / NLEN, raw copy by the whole warp
            unsigned long long src = 0; int len = 0;
            if (lane == 0) {
                br.drop(br.cnt & 7);
                const uint32_t l = br.bits(16), nl = br.bits(16);
                if ((l ^ nl) != 0xFFFFu) status = INF_BAD_BLOCK;
                len = (int)l;
                src = br.byte_pos();
                if (src + len > br.n) status = INF_SHORT; else br.seek(src + len);
            }
            status = __shfl_sync(kFull, status, 0); len = __shfl_sync(kFull, len, 0); src = __shfl_sync(kFull, src, 0);
            if (status == INF_OK && pos + len > cap) status = INF_OVERRUN;
            if (status != INF_OK) break;
            if (MODE != COUNT) {
                const uint8_t* s = zbase + src;
                for (int k = lane; k < len; k += 32) { const T v = (T)s[k]; ring[(uint32_t)(pos + k) & RM] = v; out[pos + k] = v; }
            }
            pos += len;
            filled = 0;                       // the reader jumped: refill the ring from its new position
            __syncwarp();
        } else if (btype == 3) { status = INF_BAD_BLOCK; break; }
        else {
            // ---- code tables (lane 0)
            if (lane == 0) {
                int nll = 288, nd = 30;
                if (btype == 1) {
                    for (int i = 0; i < 144; i++) M.lens[i] = 8;
                    for (int i = 144; i < 256; i++) M.lens[i] = 9;
                    for (int i = 256; i < 280; i++) M.lens[i] = 7;
                    for (int i = 280; i < 288; i++) M.lens[i] = 8;
                    for (int i = 0; i < 30; i++) M.lens[288 + i] = 5;
                } else {
                    nll = (int)br.bits(5) + 257; nd = (int)br.bits(5) + 1;
                    const int ncl = (int)br.bits(4) + 4;
                    if (nll > 286 || nd > 30) status = INF_BAD_BLOCK;
                    uint8_t cl[19];
                    for (int i = 0; i < 19; i++) cl[i] = 0;
                    for (int i = 0; i < ncl; i++) cl[kClOrd[i]] = (uint8_t)br.bits(3);
                    // the code-length code uses the d-table slots as scratch (rebuilt right after)
                    if (status == INF_OK && !build_table(cl, 19, M.cnt_d, M.sorted_d, M.fast_d)) status = INF_BAD_CODE;
                    int i = 0;
                    while (status == INF_OK && i < nll + nd) {
                        const int s = decode_sym(br, M.fast_d, M.cnt_d, M.sorted_d);
                        if (s < 0 || br.over) { status = INF_BAD_CODE; break; }
                        if (s < 16) { M.lens[i++] = (uint8_t)s; continue; }
                        int rep, v = 0;
                        if (s == 16) { if (i == 0) { status = INF_BAD_CODE; break; } v = M.lens[i - 1]; rep = 3 + (int)br.bits(2); }
                        else if (s == 17) rep = 3 + (int)br.bits(3);
                        else rep = 11 + (int)br.bits(7);
                        if (i + rep > nll + nd) { status = INF_BAD_CODE; break; }
                        while (rep--) M.lens[i++] = (uint8_t)v;
                    }
                    if (status == INF_OK) {          // move the distance lengths behind a fixed lit/len region of 288
                        uint8_t tmp[30];
                        for (int k = 0; k < nd; k++) tmp[k] = M.lens[nll + k];
                        for (int k = nll; k < 288; k++) M.lens[k] = 0;
                        for (int k = 0; k < 30; k++) M.lens[288 + k] = k < nd ? tmp[k] : 0;
                        if (M.lens[256] == 0) status = INF_BAD_CODE;
                    }
                }
                if (status == INF_OK && !build_table(M.lens, 288, M.cnt_ll, M.sorted_ll, M.fast_ll)) status = INF_BAD_CODE;
                if (status == INF_OK && !build_table(M.lens + 288, 30, M.cnt_d, M.sorted_d, M.fast_d)) status = INF_BAD_CODE;
            }
            status = __shfl_sync(kFull, status, 0);
            __syncwarp();
            // ---- tokens: lane 0 decodes a batch of up to 32 into shared memory, the warp executes it: sizes are scanned, literals
            //      are stored together, matches run in order (each copied by the whole warp through the ring)
            bool eob = false;
            while (status == INF_OK && !eob) {
                int ntok = 0;
                unsigned long long counted = 0;
                // keep the input ring ahead of the parser: a batch of 32 tokens eats at most 192 bytes, 256 counted ones 1536.
                // When executing, the next 64 words are requested now and land in the ring after the batch (latency hidden).
                const uint32_t w0 = __shfl_sync(kFull, br.word(), 0);
                uint32_t p0 = 0, p1 = 0; bool pf = false;
                if (MODE == COUNT || filled < w0 + 128) topup(M, br, filled, w0);
                else if (filled + 64 <= w0 + kZWords) {
                    pf = true;
                    p0 = __ldg(br.zw + min(filled + lane, br.maxw)); p1 = __ldg(br.zw + min(filled + 32 + lane, br.maxw));
                }
                if (lane == 0) {
                    const int lim = MODE == COUNT ? 256 : 32;
                    for (; ntok < lim; ntok++) {
                        const int s = decode_sym(br, M.fast_ll, M.cnt_ll, M.sorted_ll);
                        if (s < 0 || br.over) { status = INF_BAD_CODE; break; }
                        if (s < 256) { if (MODE == COUNT) counted++; else M.tok[ntok] = 0x80000000u | (uint32_t)s; continue; }
                        if (s == 256) { eob = true; break; }
                        if (s > 285) { status = INF_BAD_CODE; break; }
                        const int ls = s - 257;
                        const int len = kLenBase[ls] + (int)br.bits(kLenExtra[ls]);
                        const int ds = decode_sym(br, M.fast_d, M.cnt_d, M.sorted_d);
                        if (ds < 0 || ds > 29) { status = INF_BAD_CODE; break; }
                        const int dist = kDistBase[ds] + (int)br.bits(kDistExtra[ds]);
                        if (MODE == COUNT) { counted += (unsigned)len; if (pos + counted > cap) { status = INF_OVERRUN; break; } }
                        else M.tok[ntok] = ((uint32_t)dist << 9) | (uint32_t)len;
                    }
                }
                status = __shfl_sync(kFull, status, 0); eob = __shfl_sync(kFull, (int)eob, 0) != 0;
                if (status != INF_OK) break;
                if (MODE == COUNT) {
                    counted = __shfl_sync(kFull, counted, 0);
                    if (pos + counted > cap) { status = INF_OVERRUN; break; }
                    pos += counted;
                    continue;
                }
                ntok = __shfl_sync(kFull, ntok, 0);
                __syncwarp();
                const uint32_t t = lane < ntok ? M.tok[lane] : 0u;
                const bool lit = (t >> 31) != 0u;
                const int size = lane < ntok ? (lit ? 1 : (int)(t & 511u)) : 0;
                int incl = size;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += v; }
                const int total = __shfl_sync(kFull, incl, 31);
                if (pos + total > cap) { status = INF_OVERRUN; break; }
                const unsigned long long my = pos + (unsigned long long)(incl - size);
                if (lit) out[my] = (T)(t & 255u);
                uint32_t mm = __ballot_sync(kFull, lane < ntok && !lit);
                int prev = -1;
                // ring slots alias positions 32 KiB apart, so a literal enters the ring only after every match in front of it ran
                while (mm) {
                    const int f = __ffs(mm) - 1; mm &= mm - 1;
                    if (lit && lane > prev && lane < f) ring[(uint32_t)my & RM] = (T)(t & 255u);
                    prev = f;
                    const uint32_t tf = __shfl_sync(kFull, t, f);
                    const unsigned long long at = __shfl_sync(kFull, my, f);
                    const int len = (int)(tf & 511u), dist = (int)(tf >> 9);
                    if ((unsigned long long)dist > abs0 + at || (MODE == BYTES && (unsigned long long)dist > at)) { status = INF_BAD_DIST; break; }
                    __syncwarp();
                    for (int k0 = 0; k0 < len; k0 += 32) {
                        const int k = k0 + lane;
                        T v = 0;
                        if (k < len) {
                            // the source is the `dist` elements in front of the match, repeated when it is shorter than the match
                            const long long q = (long long)at - dist + (dist >= len ? k : k % dist);
                            v = (MODE == SYMS && q < 0) ? (T)(256 + kWin + q) : ring[(uint32_t)q & RM];
                        }
                        __syncwarp();
                        if (k < len) { ring[(uint32_t)(at + k) & RM] = v; out[at + k] = v; }
                    }
                    __syncwarp();
                }
                if (status != INF_OK) break;
                if (lit && lane > prev) ring[(uint32_t)my & RM] = (T)(t & 255u);
                if (pf) { M.zbuf[(filled + lane) & (kZWords - 1)] = p0; M.zbuf[(filled + 32 + lane) & (kZWords - 1)] = p1; filled += 64; }
                __syncwarp();
                pos += total;
            }
        }
        if (status != INF_OK) break;
        // ---- where did this block end?
        int stop = 0;
        if (lane == 0) {
            if (last) stop = 1;
            else if (!until_final) {
                const unsigned long long used = br.bits_used();
                if (used == br.n * 8ull) stop = 1;
                else if (used > br.n * 8ull) status = INF_SEG;
            }
            if (br.over) status = INF_SHORT;
        }
        status = __shfl_sync(kFull, status, 0);
        done = __shfl_sync(kFull, stop, 0) != 0;
    }
    *produced = pos;
    *final_seen = last != 0;
    return status;
}

// all lanes: check the two zlib header bytes and move the reader behind them
__device__ int zlib_header(BitReader& br, const uint8_t* z) {
    if (br.n < 2) return INF_SHORT;
    const uint32_t cmf = z[0], flg = z[1];
    br.seek(2);
    return ((cmf & 15) != 8 || ((cmf << 8) | flg) % 31 != 0 || (flg & 32)) ? INF_BAD_HEADER : INF_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------ serial inflate
__global__ void __launch_bounds__(32) k_inflate(DecPageD* __restrict__ pages, int n) {
    __shared__ InflMem M;
    extern __shared__ __align__(16) uint8_t dyn_smem[];
    const int lane = threadIdx.x;
    const int pg = blockIdx.x;
    if (pg >= n) return;
    DecPageD& P = pages[pg];
    if (P.status != 0 || P.mode == 1) return;
    BitReader br; br.init(P.z, P.zlen, M.zbuf);
    int status = zlib_header(br, P.z);
    unsigned long long produced = 0; bool fin = false;
    if (status == INF_OK) status = inflate_blocks<BYTES, uint8_t>(M, dyn_smem, br, P.z, P.filt, P.filt_len, 0ull, true, &produced, &fin);
    if (status == INF_OK && produced != P.filt_len) status = INF_SHORT;
    if (lane == 0) P.status = status;
}

// ------------------------------------------------------------------------------------------ segment-parallel inflate
__global__ void __launch_bounds__(128) k_infl_probe(const DecPageD* __restrict__ pages, DecSegD* __restrict__ segs, int nsegs) {
    __shared__ InflMem mem[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sg = blockIdx.x * 4 + warp;
    if (sg >= nsegs) return;
    DecSegD& S = segs[sg];
    const DecPageD& P = pages[S.page];
    if (P.status != 0) return;
    const bool first = sg == P.seg0, final_seg = sg == P.seg0 + P.nseg - 1;
    BitReader br; br.init(P.z + S.zoff, S.zlen, mem[warp].zbuf);
    int status = INF_OK;
    if (first) status = zlib_header(br, P.z);
    unsigned long long produced = 0; bool fin = false;
    if (status == INF_OK)
        status = inflate_blocks<COUNT, uint8_t>(mem[warp], nullptr, br, P.z + S.zoff, nullptr, P.filt_len, 0ull, false, &produced, &fin);
    // the stream's final block belongs in the last IDAT and nowhere else; its Adler-32 (4 bytes after byte alignment) ends the IDAT
    if (status == INF_OK && fin != final_seg) status = INF_SEG;
    if (lane == 0) {
        if (status == INF_OK && fin) { br.drop(br.cnt & 7); if (br.byte_pos() + 4 != br.n) status = INF_SEG; }
        S.olen = (uint32_t)produced;
        S.ok = status == INF_OK ? 1 : 0;
    }
}

__global__ void __launch_bounds__(32) k_infl_plan(DecPageD* __restrict__ pages, DecSegD* __restrict__ segs, int n) {
    const int pg = blockIdx.x, lane = threadIdx.x;
    if (pg >= n) return;
    DecPageD& P = pages[pg];
    if (P.status != 0 || P.nseg == 0) return;
    DecSegD* S = segs + P.seg0;
    int ok = 1; unsigned long long sum = 0;
    for (int i = lane; i < P.nseg; i += 32) { ok &= S[i].ok; sum += S[i].olen; }
    ok = __all_sync(kFull, ok);
#pragma unroll
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(kFull, sum, o);
    const bool par = ok && sum == P.filt_len;
    if (par) {
        unsigned long long base = 0;
        for (int i0 = 0; i0 < P.nseg; i0 += 32) {
            const int i = i0 + lane;
            const unsigned long long v = i < P.nseg ? S[i].olen : 0;
            unsigned long long incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += t; }
            if (i < P.nseg) S[i].opos = (uint32_t)(base + incl - v);
            base += __shfl_sync(kFull, incl, 31);
        }
    }
    if (lane == 0) P.mode = par ? 1 : 0;
}

__global__ void __launch_bounds__(32) k_infl_seg(DecPageD* __restrict__ pages, const DecSegD* __restrict__ segs, int nsegs) {
    __shared__ InflMem M;
    extern __shared__ __align__(16) uint8_t dyn_smem[];
    const int lane = threadIdx.x;
    const int sg = blockIdx.x;
    if (sg >= nsegs) return;
    const DecSegD& S = segs[sg];
    DecPageD& P = pages[S.page];
    if (P.status != 0 || P.mode != 1) return;
    BitReader br; br.init(P.z + S.zoff, S.zlen, M.zbuf);
    int status = INF_OK;
    if (sg == P.seg0) status = zlib_header(br, P.z);
    unsigned long long produced = 0; bool fin = false;
    if (status == INF_OK)
        status = inflate_blocks<SYMS, uint16_t>(M, reinterpret_cast<uint16_t*>(dyn_smem), br, P.z + S.zoff, P.sym + S.opos,
                                                (unsigned long long)S.olen, (unsigned long long)S.opos, false, &produced, &fin);
    if (status == INF_OK && produced != S.olen) status = INF_SHORT;
    if (lane == 0 && status != INF_OK) atomicMin(&P.status, status);
}

// The last 32 KiB of every IDAT, in stream order: symbol -> byte through the (already concrete) 32 KiB in front of the IDAT.
__global__ void __launch_bounds__(1024) k_infl_window(const DecPageD* __restrict__ pages, const DecSegD* __restrict__ segs, int n) {
    const int pg = blockIdx.x;
    if (pg >= n) return;
    const DecPageD& P = pages[pg];
    if (P.status != 0 || P.mode != 1) return;
    for (int s = 0; s < P.nseg; s++) {
        const DecSegD& S = segs[P.seg0 + s];
        const unsigned long long end = (unsigned long long)S.opos + S.olen;
        const unsigned long long lo = S.olen > (uint32_t)kWin ? end - kWin : S.opos;
        for (unsigned long long p = lo + threadIdx.x; p < end; p += 1024) {
            const uint32_t v = P.sym[p];
            P.filt[p] = v < 256u ? (uint8_t)v : __ldcg(P.filt + ((unsigned long long)S.opos - kWin + (v - 256u)));
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) k_infl_resolve(const DecBatchD b) {
    const int ck = blockIdx.x;
    const DecPageD& P = b.pages[b.chunk_page[ck]];
    if (P.status != 0 || P.mode != 1) return;
    const DecSegD* __restrict__ S = b.segs + P.seg0;
    const unsigned long long c0 = b.chunk_pos[ck];
    for (int it = 0; it < kResolveChunk / (256 * 8); it++) {
        const unsigned long long p0 = c0 + (unsigned long long)(it * 256 + threadIdx.x) * 8ull;
        if (p0 >= P.filt_len) break;
        // segment of p0: last s with opos <= p0
        int lo = 0, hi = P.nseg - 1;
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if ((unsigned long long)S[mid].opos <= p0) lo = mid; else hi = mid - 1; }
        int s = lo;
        unsigned long long s_end = (unsigned long long)S[s].opos + S[s].olen;
        const int cnt = (int)min(8ull, P.filt_len - p0);
        uint16_t v8[8];
        if (cnt == 8) { const uint4 q = *reinterpret_cast<const uint4*>(P.sym + p0); memcpy(v8, &q, 16); }
        else for (int k = 0; k < cnt; k++) v8[k] = P.sym[p0 + k];
        uint8_t o8[8]; bool all = cnt == 8;
        for (int k = 0; k < cnt; k++) {
            const unsigned long long p = p0 + k;
            while (p >= s_end && s + 1 < P.nseg) { s++; s_end = (unsigned long long)S[s].opos + S[s].olen; }
            const bool tail = p + kWin >= s_end;         // concrete already (k_infl_window)
            const uint32_t v = v8[k];
            uint8_t o = (uint8_t)v;
            if (!tail && v >= 256u) o = P.filt[(unsigned long long)S[s].opos - kWin + (v - 256u)];
            if (tail) all = false;
            o8[k] = o;
            v8[k] = tail ? 1 : 0;
        }
        if (all) { uint2 w; memcpy(&w, o8, 8); *reinterpret_cast<uint2*>(P.filt + p0) = w; }
        else for (int k = 0; k < cnt; k++) if (!v8[k]) P.filt[p0 + k] = o8[k];
    }
}

int decode_kernel_setup() {
    cudaError_t e = cudaFuncSetAttribute(k_inflate, cudaFuncAttributeMaxDynamicSharedMemorySize, kWin);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_infl_seg, cudaFuncAttributeMaxDynamicSharedMemorySize, kWin * 2);
    return e == cudaSuccess ? 0 : -1;
}

int launch_inflate(const DecBatchD& b, cudaStream_t st) {
    if (b.npages == 0) return 0;
    int launches = 0;
    if (b.nsegs) {
        k_infl_probe<<<(b.nsegs + 3) / 4, 128, 0, st>>>(b.pages, b.segs, b.nsegs);
        k_infl_plan<<<b.npages, 32, 0, st>>>(b.pages, b.segs, b.npages);
        k_infl_seg<<<b.nsegs, 32, kWin * 2, st>>>(b.pages, b.segs, b.nsegs);
        k_infl_window<<<b.npages, 1024, 0, st>>>(b.pages, b.segs, b.npages);
        if (b.nchunks) k_infl_resolve<<<b.nchunks, 256, 0, st>>>(b);
        launches += 4 + (b.nchunks ? 1 : 0);
    }
    k_inflate<<<b.npages, 32, kWin, st>>>(b.pages, b.npages);
    return launches + 1;
}

// ------------------------------------------------------------------------------------------ un-filter
namespace {

constexpr int kUfWarps = 4;
constexpr int kUfPitch = 136;             // bytes per staged row: up to 3 + 32 pixels x 4 channels, as 34 words (bank = 2 * lane + const)

struct UfMem {                            // views of one warp's staging buffers (separate __shared__ arrays: loads of `in` may pass stores to `out`)
    uint32_t (*in)[kUfPitch / 4];         // row r: the aligned words that cover its 32 pixels of this chunk
    uint8_t (*out)[kUfPitch];
    uint32_t* up;                         // the same for the last row of the band above
};

// aligned word `k` of the run that starts at byte address a (a itself may be misaligned); 0 outside [lo, hi)
__device__ __forceinline__ uint32_t word_at(const uint8_t* a, int k, const uint8_t* lo, const uint8_t* hi, bool l2) {
    const uint8_t* w = a - ((uintptr_t)a & 3) + 4 * k;
    if (w < lo || w + 4 > hi) return 0u;
    return l2 ? __ldcg(reinterpret_cast<const uint32_t*>(w)) : __ldg(reinterpret_cast<const uint32_t*>(w));
}

template <int BPP>
__device__ void unfilter_band(const UfMem M, const DecPageD& P, int band, uint32_t* __restrict__ flags /* of this page */, int* bad, bool nowait) {
    const int lane = threadIdx.x & 31;
    const int W = P.w, nb = W * BPP, H = P.h;
    const int y0 = band * 32, y = y0 + lane;
    const bool row_ok = y < H;
    const uint8_t* __restrict__ F = P.filt;
    uint8_t* __restrict__ X = P.pix;
    const uint8_t* Fend = F + ((P.filt_len + 3ull) & ~3ull);                      // both buffers are 256-byte aligned with slack behind
    const uint8_t* Xend = X + (((unsigned long long)nb * H + 3ull) & ~3ull);
    int ft = row_ok ? F[(unsigned long long)y * (nb + 1)] : 0;
    if (ft > 4) { *bad = 1; ft = 0; }
    const int nchunks = (W + 31 + 31) / 32;
    constexpr int NW = (3 + 32 * BPP + 3) / 4;                                    // words per staged row (<= 33)
    const long long rstep = (long long)nb + 1 - BPP;                              // row r starts r pixels behind row r - 1
    uint32_t pre[32], pre_x = 0, pre_u = 0, pre_ux = 0;                           // the next chunk, in flight while this one computes

    auto fetch = [&](int j) {
        if (band > 0) {                       // the band above must have finished the pixels lane 0 will read in chunk j
            const uint32_t need = (uint32_t)min(j + 2, nchunks);
            if (lane == 0) {
                const volatile uint32_t* f = flags + band - 1;
                while (!nowait && *f < need) __nanosleep(32);
                __threadfence();
            }
            __syncwarp();
            const uint8_t* u = X + (unsigned long long)(y0 - 1) * nb + (long long)32 * j * BPP;
            pre_u = word_at(u, lane, X, Xend, true);
            if (BPP == 4 && lane == 0) pre_ux = word_at(u, 32, X, Xend, true);
        }
        const uint8_t* a = F + (unsigned long long)y0 * (nb + 1) + 1 + (long long)32 * j * BPP;
#pragma unroll
        for (int r = 0; r < 32; r++) {
            pre[r] = (y0 + r < H && lane < NW) ? word_at(a, lane, F, Fend, false) : 0u;
            if (BPP == 4 && lane == r) pre_x = (y0 + r < H) ? word_at(a, 32, F, Fend, false) : 0u;   // word 32 of row r goes through lane r
            a += rstep;
        }
    };
    auto stage = [&]() {
        M.up[lane] = pre_u;
        if (BPP == 4 && lane == 0) M.up[32] = pre_ux;
#pragma unroll
        for (int r = 0; r < 32; r++) M.in[r][lane] = pre[r];
        if (BPP == 4) M.in[lane][32] = pre_x;
        __syncwarp();
    };

    fetch(0);
    stage();
    uint32_t cur = 0, bprev = 0;              // packed channels: my pixel x-1, the pixel above x-1
    const int ka = ft == 1 ? -1 : 0, kb = ft == 2 ? -1 : 0, kavg = ft == 3 ? -1 : 0, kp = ft == 4 ? -1 : 0;
    for (int j = 0; j < nchunks; j++) {
        if (j + 1 < nchunks) fetch(j + 1);
        const uint8_t* inb = reinterpret_cast<const uint8_t*>(M.in[lane]) +
                             ((uintptr_t)(F + (unsigned long long)y * (nb + 1) + 1 + (long long)(32 * j - lane) * BPP) & 3);
        // lane s holds pixel s of the row above (packed), handed to lane 0 by a broadcast at step s
        uint32_t upv = 0;
        if (band > 0) {
            const uint8_t* upb = reinterpret_cast<const uint8_t*>(M.up) + ((uintptr_t)(X + (unsigned long long)(y0 - 1) * nb + (long long)32 * j * BPP) & 3);
#pragma unroll
            for (int ch = 0; ch < BPP; ch++) upv |= (uint32_t)upb[lane * BPP + ch] << (8 * ch);
        }
        uint8_t* outb = M.out[lane];
#pragma unroll 8
        for (int s = 0; s < 32; s++) {
            const int x = 32 * j + s - lane;
            const uint32_t bs = __shfl_up_sync(kFull, cur, 1), b0 = __shfl_sync(kFull, upv, s);
            const uint32_t b = lane == 0 ? b0 : bs;
            uint32_t o = 0;
#pragma unroll
            for (int ch = 0; ch < BPP; ch++) {                // branch-free: the filter type differs from lane to lane
                const int a = (cur >> (8 * ch)) & 255, bb = (b >> (8 * ch)) & 255, c = (bprev >> (8 * ch)) & 255;
                const int pa = abs(bb - c), pb = abs(a - c), pc = abs(a + bb - 2 * c);
                const int t = pb <= pc ? bb : c;
                const int paeth = (pa <= pb) & (pa <= pc) ? a : t;
                const int pred = (a & ka) | (bb & kb) | (((a + bb) >> 1) & kavg) | (paeth & kp);
                const int v = ((int)inb[s * BPP + ch] + pred) & 255;
                outb[s * BPP + ch] = (uint8_t)v;
                o |= (uint32_t)v << (8 * ch);
            }
            cur = x < 0 ? 0u : o;                             // not started: left and upper-left of pixel 0 are 0
            bprev = x < 0 ? 0u : b;
        }
        __syncwarp();
        for (int r = 0; r < 32; r++) {
            if (y0 + r >= H) break;
            uint8_t* dst = X + (unsigned long long)(y0 + r) * nb;
            const int g0 = (32 * j - r) * BPP;
#pragma unroll
            for (int i = 0; i < BPP; i++) {
                const int bi = lane + 32 * i, gb = g0 + bi;
                if (gb >= 0 && gb < nb) dst[gb] = M.out[r][bi];
            }
        }
        __syncwarp();
        if (lane == 0) { __threadfence(); *(volatile uint32_t*)(flags + band) = (uint32_t)(j + 1); }
        if (j + 1 < nchunks) stage();
    }
}

}  // namespace

__global__ void __launch_bounds__(kUfWarps * 32) k_unfilter(const DecBatchD b) {
    __shared__ uint32_t s_in[kUfWarps][32][kUfPitch / 4];
    __shared__ uint8_t s_out[kUfWarps][32][kUfPitch];
    __shared__ uint32_t s_up[kUfWarps][kUfPitch / 4];
    __shared__ uint32_t ticket;
    if (threadIdx.x == 0) ticket = atomicAdd(b.counters, 1u);
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    const UfMem mem_w{s_in[warp], s_out[warp], s_up[warp]};
    const int g = (int)ticket * kUfWarps + warp;          // global band number, in launch order
    if (g >= b.nbands) return;
    int lo = 0, hi = b.npages - 1;                        // page of band g: last page with band0 <= g (pages without bands repeat band0)
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (b.pages[mid].band0 <= g) lo = mid; else hi = mid - 1; }
    DecPageD& P = b.pages[lo];
    const int band = g - P.band0;
    uint32_t* flags = b.band_flag + P.band0;
    int bad = 0;
    if (P.status != 0) {                                  // a skipped band still releases the bands waiting on it
        if ((threadIdx.x & 31) == 0) *(volatile uint32_t*)(flags + band) = 0xffffffffu;
    } else {
        switch (P.c) {
            case 1: unfilter_band<1>(mem_w, P, band, flags, &bad, b.dbg_nowait != 0); break;
            case 2: unfilter_band<2>(mem_w, P, band, flags, &bad, b.dbg_nowait != 0); break;
            case 3: unfilter_band<3>(mem_w, P, band, flags, &bad, b.dbg_nowait != 0); break;
            default: unfilter_band<4>(mem_w, P, band, flags, &bad, b.dbg_nowait != 0); break;
        }
    }
    if (__any_sync(kFull, bad) && (threadIdx.x & 31) == 0) atomicMin(&P.status, (int)INF_BAD_FILTER);
}

int launch_unfilter(const DecBatchD& b, cudaStream_t st) {
    if (b.nbands == 0) return 0;
    k_unfilter<<<(b.nbands + kUfWarps - 1) / kUfWarps, kUfWarps * 32, 0, st>>>(b);
    return 1;
}

}  // namespace vcp
