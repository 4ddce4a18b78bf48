// png_decode.cu — zlib inflate and PNG un-filtering on the GPU (the reverse of deflate_*.cu / png_filter.cu).
//
// Replaces what Pillow runs for Image.open(png).load(): zlib inflate() driven by libImaging/ZipDecode.c and its per-row
// un-filter (None/Sub/Up/Avg/Paeth).  SURVEY.md §8 f-4: re-reading the reference's images/page_###.png cache
// (backend/README.md:238-243) and validating this library's own PNGs at speed.  Restated for tests by
// oracle/restate.py:png_unfilter and Python's zlib.
//
// Inflate.  Parsing a deflate stream is serial (a code starts where the last one ended); *executing* its tokens is not, once it is
// known where in the output each stretch of tokens lands.  So the stream is parsed twice, and the first parse is itself split
// wherever a deflate block can be shown to begin:
//   k_infl_scan1/2 every bit position of the stream is tried as a dynamic-Huffman block header (complete code-length code, complete
//                  literal/length code with an end-of-block symbol, complete or one-symbol distance code: random bits do not pass).
//   k_infl_sort    the found headers and the IDAT starts (known to the host) are a page's *parse units*, sorted by bit position.
//   k_infl_probe   one warp per parse unit, parse only.  The warp counts output bytes until a block ends exactly on the first bit of
//                  another parse unit, or the final block ends.  Every 64 KiB of output it writes a checkpoint: bit position, bit
//                  position of the enclosing block header, output offset.  The first unit (bit 16 of the zlib stream) really does
//                  begin a block, so its result is true, and so is the result of the unit its parse stopped in front of, and so on: a
//                  chain of true results (k_infl_plan walks it).  zlib starts a block every 32 Ki symbols, this library one per IDAT:
//                  either way a page is parsed by many warps.  A unit that is not a block start (a false positive of the scan, an IDAT
//                  that Pillow cut mid-block) produces garbage that no chain reaches; a block the scan cannot see (stored, fixed
//                  Huffman) is parsed by the warp of the block in front of it.
//   k_infl_plan    one warp per page: walk the chain, give every interval between checkpoints its output offset, in order.
//   k_infl_exec    one warp per interval: rebuild the block's code tables, seek to the checkpoint, parse again and execute the tokens.
//                  A match may reach up to 32 KiB in front of the interval, into output another warp is still producing, so the
//                  warp works on 16-bit symbols: 0..255 = a byte, 256 + j = "byte j of the 32 KiB in front of this interval".
//                  Copies move symbols like bytes.  No shared-memory window: at ~8 KB of shared memory per warp 27 warps fit an SM,
//                  and their number, not the latency of one match through L2, sets the pace.
//   k_infl_window  one CTA per page walks the intervals in order and makes the last 32 KiB of each concrete (each step a 32 Ki-wide
//                  gather through the previous, already concrete, window).
//   k_infl_resolve every other position in parallel: symbol -> byte through the window in front of its interval.
// Un-filter.  Pixel (x, y) needs (x-1, y), (x, y-1), (x-1, y-1): a wavefront.
//   k_unfilter     one warp per band of 32 rows, lane l on row y0 + l running l pixels behind lane l - 1, so the pixel above
//                  arrives by one shuffle per step; rows are staged through shared memory 32 pixels at a time with coalesced
//                  loads and stores.  Bands of a page run concurrently three chunks apart: lane 0 reads the last row of the band
//                  above from global memory once that band's progress flag (release / acquire) covers it.  CTAs take their band
//                  range from a ticket, so a band only ever waits on bands that are already running.
// Every loop is bounded by the stream / output length: malformed input ends in a status, never in a hang.
#include "vcp_internal.cuh"
#include <type_traits>
#include <atomic>
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace vcp {

namespace {

constexpr unsigned kFull = 0xffffffffu;
#ifndef VCP_FAST_LL
#define VCP_FAST_LL 10
#endif
#ifndef VCP_FAST_D
#define VCP_FAST_D 9
#endif
constexpr int kFastLL = VCP_FAST_LL, kFastD = VCP_FAST_D;   // index bits of the lit/len and distance lookup tables
constexpr int kWin = 32768;
constexpr int kZWords = 256;              // compressed-input ring (words) the warp keeps filled ahead of lane 0's parser
constexpr int kResolveChunk = 32768;      // positions per CTA of k_infl_resolve (keep in step with api.cu)
constexpr uint32_t kCkpt = 65536;         // output bytes between checkpoints (keep in step with api.cu)

enum { INF_OK = 0, INF_BAD_HEADER = -101, INF_BAD_BLOCK = -102, INF_BAD_CODE = -103, INF_OVERRUN = -104, INF_BAD_DIST = -105,
       INF_SHORT = -106, INF_BAD_FILTER = -107, INF_SEG = -108 };

// Table entry: value << 16 | kind << 8 | extra bits << 4 | code length (0 = not in the fast table).
enum { K_LIT = 0, K_LEN = 1, K_EOB = 2, K_BAD = 3 };

struct InflMem {
    uint32_t fast_ll[1 << kFastLL];
    uint32_t fast_d[1 << kFastD];
    uint32_t zbuf[kZWords];               // the compressed stream in front of the parser: word w of the stream at zbuf[w % kZWords]
    uint32_t tok[32];                     // one batch of tokens: literal = bit 31 | byte, match = distance << 9 | length
    uint16_t sorted_ll[288], sorted_d[32];
    uint16_t cnt_ll[16], cnt_d[16];
    uint16_t offs[16], fcode[16];         // table building: first index / first code of every length
    uint8_t lens[384];                    // [0, 288) lit/len, [288, 318) distance code lengths; scratch while a header is read
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
    uint32_t limw;                               // a refill from a word beyond this one means the parse ran off the stream
    __device__ void init(const uint8_t* z, unsigned long long len, const uint32_t* ring) {
        mis = (int)((uintptr_t)z & 3); zw = reinterpret_cast<const uint32_t*>(z - mis); zs = ring;
        n = len; pos = 0; buf = 0; cnt = 0; over = false;
        maxw = (uint32_t)((len + mis + 60) >> 2);
        limw = (uint32_t)((len + mis + 12) >> 2);
        nextw = __ldg(zw);
    }
    __device__ uint32_t word() const { return (uint32_t)((pos + mis) >> 2); }
    __device__ __forceinline__ void refill() {   // call with cnt <= 32
        const uint32_t w = word();
        const uint32_t hi = zs[(w + 1) & (kZWords - 1)];
        const uint32_t v = __funnelshift_r(nextw, hi, 8 * (int)((pos + mis) & 3));   // bytes [pos, pos + 4)
        nextw = hi;
        buf |= (unsigned long long)v << cnt; cnt += 32; pos += 4;
        if (w > limw) over = true;
    }
    __device__ __forceinline__ void need33() { if (cnt <= 32) refill(); }            // afterwards at least 33 bits are buffered
    __device__ __forceinline__ uint32_t peek(int k) { return (uint32_t)buf & ((1u << k) - 1u); }   // k <= 16
    __device__ __forceinline__ void drop(int k) { buf >>= k; cnt -= k; }
    __device__ uint32_t bits(int k) { need33(); const uint32_t v = peek(k); drop(k); return v; }   // k <= 16
    __device__ unsigned long long bits_used() const { return pos * 8ull - (unsigned long long)cnt; }
    __device__ unsigned long long byte_pos() const { return pos - (unsigned long long)(cnt >> 3); }   // first byte not consumed (cnt % 8 == 0)
    __device__ void seek(unsigned long long p) { pos = p; buf = 0; cnt = 0; nextw = __ldg(zw + min((uint32_t)((p + mis) >> 2), maxw)); }
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

// Whole warp (every lane keeps an identical reader): continue at bit `bit` of the stream.
__device__ void seek_bit(InflMem& M, BitReader& br, uint32_t& filled, unsigned long long bit) {
    br.seek(bit >> 3);
    filled = 0;
    topup(M, br, filled, br.word());
    br.refill();
    br.drop((int)(bit & 7));
}

__constant__ uint16_t kLenBase[29] = {3,4,5,6,7,8,9,10,11,13,15,17,19,23,27,31,35,43,51,59,67,83,99,115,131,163,195,227,258};
__constant__ uint8_t kLenExtra[29] = {0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0};
__constant__ uint16_t kDistBase[30] = {1,2,3,4,5,7,9,13,17,25,33,49,65,97,129,193,257,385,513,769,1025,1537,2049,3073,4097,6145,8193,12289,16385,24577};
__constant__ uint8_t kDistExtra[30] = {0,0,0,0,1,1,2,2,3,3,4,4,5,5,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13};
__constant__ uint8_t kClOrd[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

enum { T_LL = 0, T_D = 1, T_CL = 2 };

// what a decoded symbol means, in table-entry form (without the code length)
template <int KIND>
__device__ __forceinline__ uint32_t sym_entry(int s) {
    if (KIND == T_CL) return (uint32_t)s << 16;
    if (KIND == T_D) return s < 30 ? ((uint32_t)kDistBase[s] << 16) | ((uint32_t)kDistExtra[s] << 4) : (uint32_t)K_BAD << 8;
    if (s < 256) return (uint32_t)s << 16;
    if (s == 256) return (uint32_t)K_EOB << 8;
    if (s > 285) return (uint32_t)K_BAD << 8;
    return ((uint32_t)kLenBase[s - 257] << 16) | ((uint32_t)K_LEN << 8) | ((uint32_t)kLenExtra[s - 257] << 4);
}

// Whole warp: canonical decode structures for n symbols with code lengths L.  Lane 0 counts and sorts, all lanes fill the lookup
// table (one symbol per lane at a time, every replica of its code).  Returns false for an over-subscribed code.
// `strict` applies zlib's inftrees.c rule to a code sent in a dynamic header: an incomplete code is an error too, unless it is a
// literal/length or distance code with a single 1-bit symbol (or no symbol at all); the fixed codes of BTYPE 1 are taken as given.
template <int KIND>
__device__ bool build_table(InflMem& M, const uint8_t* L, int n, uint16_t* cnt, uint16_t* sorted, uint32_t* fast, int fastbits, bool strict) {
    const int lane = threadIdx.x & 31;
    int ok = 1, total = 0;
    if (lane == 0) {
        for (int i = 0; i < 16; i++) cnt[i] = 0;
        for (int i = 0; i < n; i++) cnt[L[i]]++;
        cnt[0] = 0;
        int left = 1, maxl = 0;
        for (int l = 1; l < 16; l++) { left <<= 1; left -= cnt[l]; if (left < 0) ok = 0; if (cnt[l]) maxl = l; }
        if (strict && ok && left > 0 && maxl != 0 && (KIND == T_CL || maxl != 1)) ok = 0;
        uint16_t o[16]; o[1] = 0; M.offs[1] = 0; M.fcode[1] = 0;
        for (int l = 1; l < 15; l++) { o[l + 1] = o[l] + cnt[l]; M.offs[l + 1] = o[l + 1]; M.fcode[l + 1] = (uint16_t)((M.fcode[l] + cnt[l]) << 1); }
        if (ok) for (int i = 0; i < n; i++) if (L[i]) sorted[o[L[i]]++] = (uint16_t)i;
        total = o[15] + 0;                    // o[15] was advanced past the 15-bit codes: number of coded symbols
    }
    ok = __shfl_sync(kFull, ok, 0); total = __shfl_sync(kFull, total, 0);
    for (int i = lane; i < (1 << fastbits); i += 32) fast[i] = 0;
    __syncwarp();
    if (!ok) return false;
    for (int idx = lane; idx < total; idx += 32) {
        const int s = sorted[idx], l = L[s];
        if (l > fastbits) continue;
        const uint32_t code = (uint32_t)M.fcode[l] + (uint32_t)(idx - M.offs[l]);
        const uint32_t r = __brev(code) >> (32 - l);                       // the stream is LSB first
        const uint32_t e = sym_entry<KIND>(s) | (uint32_t)l;
        for (uint32_t k = r; k < (1u << fastbits); k += 1u << l) fast[k] = e;
    }
    __syncwarp();
    return true;
}

// one code (lane 0; at least 15 bits buffered): table entry including its length, K_BAD for a code that does not exist
template <int KIND>
__device__ __forceinline__ uint32_t decode_entry(BitReader& br, const uint32_t* fast, int fastbits, const uint16_t* cnt, const uint16_t* sorted) {
    const uint32_t e = fast[br.peek(fastbits)];
    if (e & 15u) return e;
    int code = 0, first = 0, index = 0;
    unsigned long long b = br.buf;
    for (int l = 1; l <= 15; l++) {
        code |= (int)(b & 1ull); b >>= 1;
        const int c = cnt[l];
        if (code - c < first) return sym_entry<KIND>(sorted[index + (code - first)]) | (uint32_t)l;
        index += c; first += c; first <<= 1; code <<= 1;
    }
    return ((uint32_t)K_BAD << 8) | 1u;
}

// Whole warp: the 3 header bits of a block and, for Huffman blocks, its code tables.  Stored blocks: *stored_len / *stored_src are
// set and the reader is moved behind the raw bytes.  Every lane returns the same status.
__device__ int block_header(InflMem& M, BitReader& br, uint32_t& filled, int* last, int* btype, int* stored_len, unsigned long long* stored_src) {
    const int lane = threadIdx.x & 31;
    topup(M, br, filled, __shfl_sync(kFull, br.word(), 0));       // a block header is at most ~330 bytes
    int status = INF_OK, bt = 0, la = 0;
    if (lane == 0) { la = (int)br.bits(1); bt = (int)br.bits(2); if (br.over) status = INF_SHORT; }
    status = __shfl_sync(kFull, status, 0); la = __shfl_sync(kFull, la, 0); bt = __shfl_sync(kFull, bt, 0);
    *last = la; *btype = bt;
    if (status != INF_OK) return status;
    if (bt == 3) return INF_BAD_BLOCK;
    if (bt == 0) {
        unsigned long long src = 0; int len = 0;
        if (lane == 0) {
            br.drop(br.cnt & 7);
            const uint32_t l = br.bits(16), nl = br.bits(16);
            if ((l ^ nl) != 0xFFFFu) status = INF_BAD_BLOCK;
            len = (int)l;
            src = br.byte_pos();
            if (src + len > br.n) status = INF_SHORT; else br.seek(src + len);
        }
        status = __shfl_sync(kFull, status, 0);
        *stored_len = __shfl_sync(kFull, len, 0); *stored_src = __shfl_sync(kFull, src, 0);
        filled = 0;                           // the reader jumped: the ring refills from its new position
        return status;
    }
    int nll = 288, nd = 30;
    if (bt == 1) {
        for (int i = lane; i < 288; i += 32) M.lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
        if (lane < 30) M.lens[288 + lane] = 5;
        __syncwarp();
    } else {
        int ncl = 0;
        if (lane == 0) {
            nll = (int)br.bits(5) + 257; nd = (int)br.bits(5) + 1; ncl = (int)br.bits(4) + 4;
            if (nll > 286 || nd > 30) status = INF_BAD_BLOCK;
            for (int i = 0; i < 19; i++) M.lens[i] = 0;
            for (int i = 0; i < ncl; i++) M.lens[kClOrd[i]] = (uint8_t)br.bits(3);
        }
        status = __shfl_sync(kFull, status, 0); nll = __shfl_sync(kFull, nll, 0); nd = __shfl_sync(kFull, nd, 0);
        if (status != INF_OK) return status;
        __syncwarp();
        // the code-length code borrows the distance table's slots (rebuilt right after)
        if (!build_table<T_CL>(M, M.lens, 19, M.cnt_d, M.sorted_d, M.fast_d, 7, true)) return INF_BAD_CODE;
        if (lane == 0) {
            uint8_t* LL = M.lens + 32;        // decoded behind the 19 code-length-code lengths, moved into place below
            int i = 0;
            while (status == INF_OK && i < nll + nd) {
                br.need33();
                const uint32_t e = decode_entry<T_CL>(br, M.fast_d, 7, M.cnt_d, M.sorted_d);
                if (((e >> 8) & 3u) == K_BAD || br.over) { status = INF_BAD_CODE; break; }
                br.drop((int)(e & 15u));
                const int s = (int)(e >> 16);
                if (s < 16) { LL[i++] = (uint8_t)s; continue; }
                int rep, v = 0;
                if (s == 16) { if (i == 0) { status = INF_BAD_CODE; break; } v = LL[i - 1]; rep = 3 + (int)br.peek(2); br.drop(2); }
                else if (s == 17) { rep = 3 + (int)br.peek(3); br.drop(3); }
                else { rep = 11 + (int)br.peek(7); br.drop(7); }
                if (i + rep > nll + nd) { status = INF_BAD_CODE; break; }
                while (rep--) LL[i++] = (uint8_t)v;
            }
            if (status == INF_OK) {           // lit/len lengths to [0, 288), distance lengths to [288, 318); in-place moves run downwards / via a copy
                uint8_t tmp[30];
                for (int k = 0; k < nd; k++) tmp[k] = LL[nll + k];
                for (int k = 0; k < nll; k++) M.lens[k] = LL[k];
                for (int k = nll; k < 288; k++) M.lens[k] = 0;
                for (int k = 0; k < 30; k++) M.lens[288 + k] = k < nd ? tmp[k] : 0;
                if (M.lens[256] == 0) status = INF_BAD_CODE;
            }
        }
        status = __shfl_sync(kFull, status, 0);
        if (status != INF_OK) return status;
        __syncwarp();
    }
    if (!build_table<T_LL>(M, M.lens, 288, M.cnt_ll, M.sorted_ll, M.fast_ll, kFastLL, bt == 2)) return INF_BAD_CODE;
    if (!build_table<T_D>(M, M.lens + 288, 30, M.cnt_d, M.sorted_d, M.fast_d, kFastD, bt == 2)) return INF_BAD_CODE;
    return INF_OK;
}


// ---- warp-parallel token parse.  Where a token starts is only known once the one in front of it has been decoded, but *what* a
// token would be if it started at a given bit can be looked up for 32 consecutive bits at once: lane l decodes a whole token
// (literal, or length + distance with their extra bits) as if it began at bit B + l.  The true tokens are then found by following
// the chain 0 -> tb(0) -> ... with one shuffle per token, instead of a dependent table lookup, shifts and branches per token.
enum { SP_EOB = 1 << 15, SP_BAD = 1 << 16, SP_SLOW = 1 << 17 };    // info = bits | size << 6 | flags

template <bool ALLOW_SLOW>
__device__ __forceinline__ void spec_decode(const InflMem& M, uint32_t w0, uint32_t w1, uint32_t* info, uint32_t* tok) {
    uint32_t e = M.fast_ll[w0 & ((1u << kFastLL) - 1u)];
    if ((e & 15u) == 0u) {
        if (!ALLOW_SLOW) { *info = SP_SLOW; return; }
        int code = 0, first = 0, index = 0; uint32_t b = w0; e = ((uint32_t)K_BAD << 8) | 1u;
        for (int l = 1; l <= 15; l++) {
            code |= (int)(b & 1u); b >>= 1;
            const int c = M.cnt_ll[l];
            if (code - c < first) { e = sym_entry<T_LL>(M.sorted_ll[index + (code - first)]) | (uint32_t)l; break; }
            index += c; first += c; first <<= 1; code <<= 1;
        }
    }
    const uint32_t nb = e & 15u, kind = (e >> 8) & 3u, val = e >> 16;
    if (kind == K_LIT) { *info = nb | (1u << 6); *tok = 0x80000000u | val; return; }
    if (kind != K_LEN) { *info = nb | (kind == K_EOB ? SP_EOB : SP_BAD); return; }
    const uint32_t xl = (e >> 4) & 15u;
    const uint32_t len = val + ((w0 >> nb) & ((1u << xl) - 1u));
    const uint32_t c1 = nb + xl;                                             // <= 20
    const uint32_t w2 = __funnelshift_r(w0, w1, c1);
    uint32_t d = M.fast_d[w2 & ((1u << kFastD) - 1u)];
    if ((d & 15u) == 0u) {
        if (!ALLOW_SLOW) { *info = SP_SLOW; return; }
        int code = 0, first = 0, index = 0; uint32_t b = w2; d = ((uint32_t)K_BAD << 8) | 1u;
        for (int l = 1; l <= 15; l++) {
            code |= (int)(b & 1u); b >>= 1;
            const int c = M.cnt_d[l];
            if (code - c < first) { d = sym_entry<T_D>(M.sorted_d[index + (code - first)]) | (uint32_t)l; break; }
            index += c; first += c; first <<= 1; code <<= 1;
        }
    }
    if (((d >> 8) & 3u) == K_BAD) { *info = SP_BAD; return; }
    const uint32_t nbd = d & 15u, xd = (d >> 4) & 15u;
    const uint32_t dist = (d >> 16) + ((w2 >> nbd) & ((1u << xd) - 1u));     // nbd + xd <= 28
    *info = (c1 + nbd + xd) | (len << 6);
    *tok = (dist << 9) | len;
}

// the 64 stream bits that start at bit position `bit` (relative to the first byte of the stream), from the shared-memory ring
__device__ __forceinline__ void ring_window(const InflMem& M, const BitReader& br, unsigned long long bit, uint32_t* w0, uint32_t* w1) {
    const unsigned long long a = bit + 8ull * (unsigned)br.mis;
    const uint32_t i = (uint32_t)(a >> 5); const int sh = (int)(a & 31);
    const uint32_t x0 = M.zbuf[i & (kZWords - 1)], x1 = M.zbuf[(i + 1) & (kZWords - 1)], x2 = M.zbuf[(i + 2) & (kZWords - 1)];
    *w0 = __funnelshift_r(x0, x1, sh); *w1 = __funnelshift_r(x1, x2, sh);
}

}  // namespace

// ------------------------------------------------------------------------------------------ scan: where do deflate blocks start?
// A parse can begin wherever a deflate block begins, and the header of a dynamic-Huffman block (the only kind an encoder emits for
// page-sized data) is recognisable: 3 + 14 fixed bits, a code-length code that must be complete, and ~300 run-length coded code
// lengths that must form a complete literal/length code containing the end-of-block symbol and a complete (or one-symbol)
// distance code.  Random bits pass that with probability ~0, so every bit position of the stream is simply tried:
//   k_infl_scan1   thread per bit position: the fixed fields and the Kraft sum of the code-length code (cheap, ~0.4 % pass)
//   k_infl_scan2   thread per survivor: decode the code lengths, tracking the two Kraft sums (leaves at the first overflow)
//   k_infl_sort    CTA per page: the IDAT starts (from the host) and the found headers, sorted, become the page's parse units
// A false positive only costs a wasted parse (no chain reaches it); a missed block (stored / fixed Huffman) is parsed by the warp of
// the block in front of it.
constexpr int kScanBits = 8192;           // bit positions per CTA of k_infl_scan1 (keep in step with api.cu)

// 32 stream bits starting at bit `bit` of the byte stream z (z has >= 64 bytes of slack behind its n bytes)
__device__ __forceinline__ uint32_t bits_at(const uint8_t* __restrict__ z, unsigned long long bit) {
    const uintptr_t a = (uintptr_t)z + (uintptr_t)(bit >> 3);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
    const int sh = (int)((a & 3) * 8 + (bit & 7));                             // 0..31
    return __funnelshift_r(__ldg(w), __ldg(w + 1), sh);
}

__global__ void __launch_bounds__(256) k_infl_scan1(const DecBatchD b) {
    const int ck = blockIdx.x;
    DecPageD& P = b.pages[b.scan_page[ck]];
    if (P.status != 0) return;
    const unsigned long long nbits = P.zlen * 8ull;
    __shared__ uint32_t s_n, s_base, s_list[512];                              // survivors of this CTA: one global atomic for all of them
    __shared__ uint8_t s_kraft[512];                                           // Kraft sum (in 1/128) of three 3-bit code lengths
    for (int i = threadIdx.x; i < 512; i += 256) {
        uint32_t k = 0;
        for (int f = 0; f < 3; f++) { const uint32_t len = (i >> (3 * f)) & 7u; if (len) k += 128u >> len; }
        s_kraft[i] = (uint8_t)k;
    }
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    // a thread tries 32 consecutive bit positions out of one 128-bit window held in registers: first the fixed fields of all 32
    // (a bit mask), then the code-length code of those that passed — lanes loop over their set bits, not over all positions
    const unsigned long long t0 = (unsigned long long)b.scan_bit[ck] + 32ull * threadIdx.x;
    if (t0 < nbits) {
        const uintptr_t a = (uintptr_t)P.z + (uintptr_t)(t0 >> 3);              // t0 is a multiple of 32: (t0 & 7) == 0
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
        const int sh0 = (int)(a & 3) * 8;
        uint32_t x[5];
#pragma unroll
        for (int i = 0; i < 5; i++) x[i] = __ldg(wp + i);
        uint32_t w[4];                                                         // stream bits [t0 + 32 i, t0 + 32 i + 32)
#pragma unroll
        for (int i = 0; i < 4; i++) w[i] = __funnelshift_r(x[i], x[i + 1], sh0);
        uint32_t pass = 0;
#pragma unroll
        for (int o = 0; o < 32; o++) {
            const uint32_t h = __funnelshift_r(w[0], w[1], o);
            const bool ok = ((h >> 1) & 3u) == 2u && ((h >> 3) & 31u) <= 29u && ((h >> 8) & 31u) <= 29u;   // BTYPE = dynamic, HLIT, HDIST
            pass |= (ok ? 1u : 0u) << o;
        }
        while (pass) {
            const int o = __ffs(pass) - 1; pass &= pass - 1;
            const unsigned long long bit = t0 + o;
            if (bit < 17 || bit + 60 > nbits) continue;                        // bit 16 is a parse unit anyway; a header needs room
            const uint32_t h = __funnelshift_r(w[0], w[1], o);
            const int nb = 3 * ((int)((h >> 13) & 15u) + 4);                   // bits of code-length-code lengths: 12..57, from bit 17
            const int p0 = o + 17, p1 = o + 47;
            uint32_t c0 = p0 < 32 ? __funnelshift_r(w[0], w[1], p0) : __funnelshift_r(w[1], w[2], p0 - 32);
            uint32_t c1 = p1 < 64 ? __funnelshift_r(w[1], w[2], p1 - 32) : __funnelshift_r(w[2], w[3], p1 - 64);
            c0 &= nb >= 30 ? 0x3FFFFFFFu : (1u << nb) - 1u;
            c1 &= nb > 30 ? (1u << (nb - 30)) - 1u : 0u;
            // complete iff the Kraft sum is exactly 1 (128 / 128)
            const uint32_t kraft = s_kraft[c0 & 511u] + s_kraft[(c0 >> 9) & 511u] + s_kraft[(c0 >> 18) & 511u] + s_kraft[c0 >> 27] +
                                   s_kraft[c1 & 511u] + s_kraft[(c1 >> 9) & 511u] + s_kraft[(c1 >> 18) & 511u];
            if (kraft != 128u) continue;
            const uint32_t q = atomicAdd(&s_n, 1u);
            if (q < 512u) s_list[q] = (uint32_t)bit;
            else { const uint32_t idx = atomicAdd(&P.nsurv, 1u); if (idx < (uint32_t)P.surv_cap) b.surv[P.surv0 + idx] = (uint32_t)bit; }
        }
    }
    __syncthreads();
    const uint32_t n = min(s_n, 512u);
    if (threadIdx.x == 0 && n) s_base = atomicAdd(&P.nsurv, n);
    __syncthreads();
    for (uint32_t q = threadIdx.x; q < n; q += 256) { const uint32_t idx = s_base + q; if (idx < (uint32_t)P.surv_cap) b.surv[P.surv0 + idx] = s_list[q]; }
}

__global__ void __launch_bounds__(128) k_infl_scan2(const DecBatchD b) {
    const int g = blockIdx.x * 128 + threadIdx.x;
    if (g >= b.surv_total) return;
    int lo = 0, hi = b.npages - 1;                        // page of survivor slot g: last page with surv0 <= g
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (b.pages[mid].surv0 <= g) lo = mid; else hi = mid - 1; }
    DecPageD& P = b.pages[lo];
    if (P.status != 0 || (uint32_t)(g - P.surv0) >= min(P.nsurv, (uint32_t)P.surv_cap)) return;
    const uint8_t* __restrict__ z = P.z;
    const unsigned long long nbits = P.zlen * 8ull;
    const unsigned long long bit0 = b.surv[g];
    const uint32_t h = bits_at(z, bit0);
    const int nll = (int)((h >> 3) & 31u) + 257, nd = (int)((h >> 8) & 31u) + 1, ncl = (int)((h >> 13) & 15u) + 4;
    // the code-length code (<= 7 bits, complete): direct lookup, entry = symbol << 3 | length
    uint8_t cl[19], tab[128];
    {
        const uint32_t c0 = bits_at(z, bit0 + 17), c1 = bits_at(z, bit0 + 17 + 30);
#pragma unroll
        for (int k = 0; k < 19; k++) {
            const uint32_t len = k < 10 ? (c0 >> (3 * k)) & 7u : (c1 >> (3 * (k - 10))) & 7u;
            cl[kClOrd[k]] = k < ncl ? (uint8_t)len : (uint8_t)0;
        }
        uint32_t code = 0;
        for (int l = 1; l <= 7; l++) {
            for (int s = 0; s < 19; s++) {
                if (cl[s] != l) continue;
                const uint32_t r = __brev(code) >> (32 - l);
                for (uint32_t e = r; e < 128u; e += 1u << l) tab[e] = (uint8_t)((s << 3) | l);
                code++;
            }
            code <<= 1;
        }
    }
    unsigned long long bit = bit0 + 17 + 3ull * ncl;
    uint32_t kr_ll = 0, kr_d = 0;                         // Kraft sums in units of 2^-15
    int i = 0, prev = 0, n_d = 0; bool has_eob = false, ok = true;
    while (i < nll + nd) {
        if (bit + 14 > nbits) { ok = false; break; }
        const uint32_t w = bits_at(z, bit);
        const uint32_t e = tab[w & 127u];
        const int l = (int)(e & 7u), s = (int)(e >> 3);
        int rep = 1, v = s;
        if (s < 16) bit += l;
        else if (s == 16) { if (i == 0) { ok = false; break; } v = prev; rep = 3 + (int)((w >> l) & 3u); bit += l + 2; }
        else if (s == 17) { v = 0; rep = 3 + (int)((w >> l) & 7u); bit += l + 3; }
        else { v = 0; rep = 11 + (int)((w >> l) & 127u); bit += l + 7; }
        if (i + rep > nll + nd) { ok = false; break; }
        if (v) {
            for (int r = 0; r < rep; r++) {
                if (i + r < nll) { kr_ll += 32768u >> v; if (i + r == 256) has_eob = true; }
                else { kr_d += 32768u >> v; n_d++; }
            }
            if (kr_ll > 32768u || kr_d > 32768u) { ok = false; break; }
        }
        prev = v; i += rep;
    }
    if (!ok || !has_eob || kr_ll != 32768u || !(kr_d == 32768u || n_d <= 1)) return;
    // an IDAT start is a parse unit already
    {
        const unsigned long long* I = b.cand_bits + P.cand0;
        int a = 0, c = P.n_idat - 1;
        while (a < c) { const int mid = (a + c + 1) >> 1; if (I[mid] <= bit0) a = mid; else c = mid - 1; }
        if (I[a] == bit0) return;
    }
    const uint32_t idx = atomicAdd(&P.ncand, 1u);
    if (idx < (uint32_t)P.seg_cap) b.cand_bits[P.cand0 + idx] = bit0;
}

// The page's parse units in stream order (rank sort: the candidates are distinct).
__global__ void __launch_bounds__(256) k_infl_sort(const DecBatchD b) {
    DecPageD& P = b.pages[blockIdx.x];
    if (P.status != 0) return;
    if (threadIdx.x == 0) {
        if (P.zlen < 6) P.status = INF_SHORT;
        else { const uint32_t cmf = P.z[0], flg = P.z[1]; if ((cmf & 15) != 8 || (cmf >> 4) > 7 || ((cmf << 8) | flg) % 31 != 0 || (flg & 32)) P.status = INF_BAD_HEADER; }   // inflate(): method, window size, check bits, no preset dictionary
    }
    const int K = (int)min(P.ncand, (uint32_t)P.seg_cap);
    const unsigned long long* C = b.cand_bits + P.cand0;
    const unsigned long long* H = b.cand_hdr + P.cand0;
    for (int i = threadIdx.x; i < K; i += 256) {
        const unsigned long long v = C[i];
        int rank = 0;
        for (int j = 0; j < K; j++) rank += (C[j] < v || (C[j] == v && j < i)) ? 1 : 0;     // stable: equal start bits keep distinct slots
        DecSegD S; memset(&S, 0, sizeof S);
        S.page = blockIdx.x; S.start_bit = v; S.hdr_bit = H[i] ? H[i] : v;
        S.iv0 = (uint32_t)P.slot0 + (uint32_t)rank * (uint32_t)P.page_iv; S.iv_cap = (uint32_t)P.page_iv;
        b.segs[P.seg0 + rank] = S;
    }
    if (threadIdx.x == 0) P.nseg = K;
}

// ------------------------------------------------------------------------------------------ spec: parse units INSIDE long blocks
// The first parse is serial per parse unit, and a unit is a deflate block: zlib starts one every 32 Ki symbols, but this library's
// own PNGs (512 KiB of input per block: 400 K tokens on a photo) and other encoders' large blocks leave one warp parsing alone for
// tens of milliseconds.  Huffman-coded streams re-synchronise: a parse started at a WRONG bit falls into step with the true token
// boundaries after a few dozen tokens.  So every kSpecBits of stream one warp builds the tables of the block that (as far as the
// found headers say) contains that point, starts decoding at the point itself, and after kSpecSync bits reports the token boundary
// it has reached as a further candidate parse unit "inside block H".  Nothing is assumed about it: k_infl_probe only ends a
// unit's parse on a candidate when its own token boundary falls exactly on the candidate's start AND it is inside the same block
// (same header bit) — from there on the two parses are identical by construction.  A candidate that is not a true boundary is
// walked over and never reached by the chain, like a false positive of the header scan.
#ifndef VCP_SPEC_SYNC
#define VCP_SPEC_SYNC 6144
#endif
constexpr unsigned long long kSpecSync = VCP_SPEC_SYNC;

__global__ void __launch_bounds__(128) k_infl_spec(const DecBatchD b) {
    __shared__ InflMem mem[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x * 4 + warp;
    if (g >= b.spec_total) return;
    int plo = 0, phi = b.npages - 1;                    // page of speculative point g: last page with spec0 <= g
    while (plo < phi) { const int mid = (plo + phi + 1) >> 1; if (b.pages[mid].spec0 <= g) plo = mid; else phi = mid - 1; }
    DecPageD& P = b.pages[plo];
    const int t = g - P.spec0;
    if (P.status != 0 || t >= P.nspec || P.nseg <= 0) return;
    const unsigned long long nbits = P.zlen * 8ull, kSpecBits = P.spec_bits;                // api.cu sizes the candidate ranges with the same number
    const unsigned long long point = (unsigned long long)(t + 1) * kSpecBits;
    if (point + kSpecSync + 2048 >= nbits) return;
    const DecSegD* PS = b.segs + P.seg0;                // the block starts found so far, sorted
    int lo = 0, hi = P.nseg - 1;
    if (PS[0].start_bit > point) return;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (PS[mid].start_bit <= point) lo = mid; else hi = mid - 1; }
    if (point - PS[lo].start_bit < kSpecBits / 2) return;                                   // a unit starts close in front anyway
    if (lo + 1 < P.nseg && PS[lo + 1].start_bit < point + kSpecSync + 2048) return;         // ... or close behind
    InflMem& M = mem[warp];
    BitReader br; br.init(P.z, P.zlen, M.zbuf);
    uint32_t filled = 0;
    // The block that contains the point starts at the last found start in front of it that really is a block header.  IDAT starts
    // are in the list too, and a foreign encoder cuts IDATs in the middle of blocks (Pillow: every 64 KiB): a start that does not
    // read as a dynamic-Huffman header is stepped over.  (A wrong pick costs nothing but this point: k_infl_probe only ends a unit
    // on a candidate whose header is the block it is parsing.)
    unsigned long long H = 0;
    bool found = false;
    for (int back = 0; back < 6 && lo - back >= 0; back++) {
        H = PS[lo - back].start_bit;
        filled = 0;
        seek_bit(M, br, filled, H);
        int last = 0, btype = 0, slen = 0; unsigned long long ssrc = 0;
        if (block_header(M, br, filled, &last, &btype, &slen, &ssrc) == INF_OK && btype == 2) { found = true; break; }
        if (back == 0 && P.n_idat > 0 && lo == 0) break;                                    // the stream start is a block start: nothing in front of it
    }
    if (!found) return;
    if (__shfl_sync(kFull, br.bits_used(), 0) >= point) return;
    unsigned long long B = point;
    filled = 0;
    int resync = 0;
    while (B < point + kSpecSync) {
        topup(M, br, filled, (uint32_t)((B + 8ull * (unsigned)br.mis) >> 5));
        uint32_t w0, w1, info = 0, tok = 0;
        ring_window(M, br, B + lane, &w0, &w1);
        spec_decode<true>(M, w0, w1, &info, &tok);
        int o = 0; bool stop = false;
        while (o < 32) {
            const uint32_t inf = __shfl_sync(kFull, info, o);
            if (inf & SP_EOB) { stop = true; break; }                                        // a block ends (or seems to): leave it to the header scan
            if (inf & SP_BAD) { if (++resync > 64) stop = true; o += 1; continue; }          // not a code here: try the next bit
            o += (int)(inf & 63u);
        }
        if (stop) return;
        B += o;
    }
    if (lane == 0) {
        const uint32_t idx = atomicAdd(&P.ncand, 1u);
        if (idx < (uint32_t)P.seg_cap) { b.cand_bits[P.cand0 + idx] = B; b.cand_hdr[P.cand0 + idx] = H; }
    }
}

// ------------------------------------------------------------------------------------------ probe: parse, count, checkpoint
__global__ void __launch_bounds__(128) k_infl_probe(const DecPageD* __restrict__ pages, DecSegD* __restrict__ segs, DecIvD* __restrict__ slots,
                                                    int npages, int seg_total) {
    __shared__ InflMem mem[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sg = blockIdx.x * 4 + warp;
    if (sg >= seg_total) return;
    int plo = 0, phi = npages - 1;                      // page of parse-unit slot sg: last page with seg0 <= sg
    while (plo < phi) { const int mid = (plo + phi + 1) >> 1; if (pages[mid].seg0 <= sg) plo = mid; else phi = mid - 1; }
    const DecPageD& P = pages[plo];
    const int s_loc = sg - P.seg0;
    if (P.status != 0 || s_loc >= P.nseg) return;
    DecSegD& S = segs[sg];
    const long long t_dbg0 = clock64();
    InflMem& M = mem[warp];
    const DecSegD* PS = segs + P.seg0;                  // this page's parse units, sorted by start bit
    BitReader br; br.init(P.z, P.zlen, M.zbuf);
    uint32_t filled = 0;
    int status = INF_OK;
    const unsigned long long start_bit = S.start_bit, unit_hdr = S.hdr_bit;   // unit_hdr != start_bit: the unit starts inside a block (k_infl_spec)
    seek_bit(M, br, filled, unit_hdr);
    const unsigned long long cap = P.filt_len, nbits = P.zlen * 8ull;
    unsigned long long pos = 0;                         // output bytes so far
    // interval being built (lane 0)
    unsigned long long iv_hdr = unit_hdr, iv_start = start_bit, iv_out = 0, next_ck = kCkpt;
    uint32_t niv = 0;
    int end_seg = s_loc + 1;                            // the first parse unit whose start the parse has not passed yet
    int last = 0, next = P.nseg;
    bool done = false, first_block = true;
    // every lane: the next unit in stream order, for the token-boundary test inside blocks
    int ns = s_loc + 1;
    unsigned long long nxt_start = ns < P.nseg ? PS[ns].start_bit : ~0ull;
    auto emit = [&](unsigned long long out_now, unsigned long long hdr_bit, unsigned long long bit_now) {   // lane 0: close the interval at a token boundary
        if (out_now > iv_out) {
            if (niv < S.iv_cap) { DecIvD& I = slots[S.iv0 + niv]; I.hdr_bit = iv_hdr; I.start_bit = iv_start; I.out = (uint32_t)iv_out; I.len = (uint32_t)(out_now - iv_out); I.seg = (uint32_t)sg; I.last = 0; }
            niv++;
        }
        iv_hdr = hdr_bit; iv_start = bit_now; iv_out = out_now;
        next_ck = (out_now / kCkpt + 1) * kCkpt;
    };
    while (status == INF_OK && !done) {
        unsigned long long hdr_bit = 0;
        if (lane == 0) hdr_bit = br.bits_used();
        int btype = 0, slen = 0; unsigned long long ssrc = 0;
        status = block_header(M, br, filled, &last, &btype, &slen, &ssrc);
        if (__shfl_sync(kFull, br.bits_used(), 0) > nbits) status = INF_SHORT;   // the header runs past the end of the data (bits behind it are not the stream's)
        if (status != INF_OK) break;
        hdr_bit = __shfl_sync(kFull, hdr_bit, 0);
        const bool mid_start = first_block && unit_hdr != start_bit;
        first_block = false;
        if (mid_start && (btype == 0 || __shfl_sync(kFull, br.bits_used(), 0) > start_bit)) { status = INF_SEG; break; }   // never offered by k_infl_spec
        if (btype == 0) {
            pos += slen;
            if (pos > cap) { status = INF_OVERRUN; break; }
        } else {
            bool eob = false, unit_stop = false;
            unsigned long long B = mid_start ? start_bit : __shfl_sync(kFull, br.bits_used(), 0);      // bit position of the next token
            uint32_t p = (uint32_t)pos;                                       // cap < 2^32, a token adds <= 258: no wrap before the checks
            uint32_t ck = (uint32_t)min(__shfl_sync(kFull, next_ck, 0), 0xffffffffull);
            while (status == INF_OK && !eob && !unit_stop) {
                if (B > br.n * 8ull + 64ull) { status = INF_SHORT; break; }
                topup(M, br, filled, (uint32_t)((B + 8ull * (unsigned)br.mis) >> 5));
                for (int r = 0; r < 48 && !eob && !unit_stop && status == INF_OK; r++) {    // 48 rounds eat at most 48 * 79 bits = 119 words
                    uint32_t w0, w1, info = 0, tok = 0;
                    ring_window(M, br, B + lane, &w0, &w1);
                    spec_decode<false>(M, w0, w1, &info, &tok);
                    int o = 0;
                    const bool near_end = B + 2048ull > nbits;                // a round eats < 2048 bits: only the last ones look at the end
                    const bool near_unit = nxt_start < B + 2048ull;           // ... or at the next parse unit's start
                    while (o < 32) {
                        if (near_unit) {
                            // does a token boundary of this parse fall exactly on the start of another unit of the same block?  Then
                            // that unit's parse is this one's continuation: stop here.  Units walked over were not on a boundary.
                            const unsigned long long at = B + o;
                            while (nxt_start < at || (nxt_start == at && PS[ns].hdr_bit != hdr_bit)) { ns++; nxt_start = ns < P.nseg ? PS[ns].start_bit : ~0ull; }
                            if (nxt_start == at && at > start_bit) { unit_stop = true; break; }
                        }
                        uint32_t inf = __shfl_sync(kFull, info, o);
                        if ((inf & (SP_SLOW | SP_BAD | SP_EOB)) || near_end) {    // the common token pays one test for all of these
                            if (inf & SP_SLOW) {
                                if (lane == o) spec_decode<true>(M, w0, w1, &info, &tok);
                                inf = __shfl_sync(kFull, info, o);
                            }
                            // a code cut off by the end of the data is no code (the bits behind the data are not the stream's)
                            if (near_end && B + o + ((inf & SP_BAD) ? 48u : (inf & 63u)) > nbits) { status = INF_SHORT; break; }
                            if (inf & SP_BAD) { status = INF_BAD_CODE; break; }
                            if (inf & SP_EOB) { o += (int)(inf & 63u); eob = true; break; }
                        }
                        o += (int)(inf & 63u);
                        p += (inf >> 6) & 511u;
                        if (p >= ck) {
                            if (p > cap) { status = INF_OVERRUN; break; }
                            if (lane == 0) emit(p, hdr_bit, B + o);
                            ck = (uint32_t)min(((unsigned long long)p / kCkpt + 1) * kCkpt, 0xffffffffull);
                        }
                    }
                    B += o;
                }
            }
            if (status == INF_OK && p > cap) status = INF_OVERRUN;
            pos = p;
            if (status == INF_OK && unit_stop) { next = ns; last = 0; done = true; break; }   // (all lanes) handed over to the unit that starts here
            if (status == INF_OK) seek_bit(M, br, filled, B);                 // the header reader continues behind the end-of-block code
            if (status != INF_OK) break;
        }
        // ---- where did this block end?  (lane 0 decides)
        int stop = 0;
        if (lane == 0) {
            const unsigned long long used = br.bits_used();
            if (pos >= next_ck) emit(pos, used, used);                       // after a stored block
            if (last) stop = 1;
            else {
                // units are sorted: gallop, then bisect to the first one that starts at or behind this block's end
                int step = 1, hi2 = end_seg;
                while (hi2 < P.nseg && PS[hi2].start_bit < used) { end_seg = hi2 + 1; hi2 += step; step <<= 1; }
                hi2 = min(hi2, P.nseg);
                while (end_seg < hi2) { const int mid = (end_seg + hi2) >> 1; if (PS[mid].start_bit < used) end_seg = mid + 1; else hi2 = mid; }
                while (end_seg < P.nseg && PS[end_seg].start_bit == used && PS[end_seg].hdr_bit != used) end_seg++;   // only a unit that IS a block start
                if (end_seg < P.nseg && PS[end_seg].start_bit == used) { stop = 1; next = end_seg; }
            }
            if (br.over && status == INF_OK) status = INF_SHORT;
        }
        status = __shfl_sync(kFull, status, 0);
        done = __shfl_sync(kFull, stop, 0) != 0;
    }
    if (lane == 0) {
        // A parse that failed (or ran past the size of the image) still keeps what it produced up to there: Pillow's decoder stops
        // at the last row of the image and never looks at what follows, so k_infl_plan accepts a failed unit if the image ends in it.
        const unsigned long long used = br.bits_used();
        emit(min(pos, cap), used, used);
        if (niv > S.iv_cap) { status = INF_OVERRUN; pos = 0; niv = 0; }
        S.olen = (uint32_t)min(pos, cap); S.next = next; S.fin = status == INF_OK ? last : 0; S.niv = niv;
        S.end_bit = (status == INF_OK && last) ? used : 0ull;
        S.ok = status == INF_OK ? 1 : status;
        S.pad = (uint32_t)((clock64() - t_dbg0) >> 10);         // debug: kilo-cycles this unit's parse took (VCP_DECODE_DEBUG)
    }
}

// Walk the chain of parse units whose parse began at a true block boundary; copy their intervals, in stream order, with absolute
// offsets.  The image ends where filt_len bytes have been produced — Pillow's decoder (ZipDecode.c) stops at its last row and never
// looks at what follows — so a unit that failed or overran after that point is as good as any; the interval that reaches the end is
// clipped and marked (k_infl_exec inspects what lies behind it).  A stream whose final block ends early is kept too (valid_len).
// One CTA per page.  The walk itself is serial (a unit names its successor), so it runs on one thread over a shared-memory copy of the
// units' few fields (a tile of 1024 units at a time, loaded by the whole CTA); the intervals of the units it visited are then copied
// by all warps.  (One warp walking through global memory took 1.2 us per unit: 0.6 ms for a photo page's 500 units.)
constexpr int kPlanTile = 1024, kPlanThreads = 256;
__global__ void __launch_bounds__(kPlanThreads) k_infl_plan(DecPageD* __restrict__ pages, DecSegD* __restrict__ segs, const DecIvD* __restrict__ slots,
                                                            DecIvD* __restrict__ ivs, int n) {
    __shared__ int32_t t_ok[kPlanTile], t_next[kPlanTile];
    __shared__ uint32_t t_olen[kPlanTile], t_niv[kPlanTile], t_iv0[kPlanTile];
    __shared__ uint8_t t_fin[kPlanTile];
    __shared__ uint16_t v_idx[kPlanTile];
    __shared__ uint32_t v_opos[kPlanTile], v_ivoff[kPlanTile], v_keep[kPlanTile];
    __shared__ int sh_s, sh_nv, sh_done;
    const int pg = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (pg >= n) return;
    DecPageD& P = pages[pg];
    if (P.status != 0) return;
    DecSegD* S = segs + P.seg0;
    const unsigned long long flen = P.filt_len;
    const int nseg = P.nseg;
    // thread 0's walk state
    unsigned long long opos = 0; uint32_t niv = 0;
    int fin = 0, status = INF_OK, guard = 0, last_s = -1;
    bool complete = false;
    if (tid == 0) { sh_s = 0; sh_done = nseg <= 0; }
    __syncthreads();
    while (!sh_done) {
        const int base = sh_s, cnt = min(kPlanTile, nseg - base);
        for (int i = tid; i < cnt; i += kPlanThreads) {
            const DecSegD& U = S[base + i];
            t_ok[i] = U.ok; t_next[i] = U.next; t_olen[i] = U.olen; t_niv[i] = U.niv; t_iv0[i] = U.iv0; t_fin[i] = (uint8_t)(U.fin != 0);
        }
        __syncthreads();
        if (tid == 0) {
            int s = base, nv = 0, done = 0;
            while (s >= base && s < base + cnt) {
                if (guard++ > nseg) { done = 1; break; }
                if (nv >= cnt) break;                 // (only a chain that does not move forward gets here: the guard ends it)
                const int j = s - base;
                const int ok = t_ok[j];
                const unsigned long long avail = t_olen[j];
                const bool enough = opos + avail >= flen;
                if (ok != 1 && !enough) { status = ok < 0 ? ok : INF_SEG; done = 1; break; }
                if (niv + t_niv[j] > (uint32_t)P.iv_cap) { status = INF_OVERRUN; done = 1; break; }
                uint32_t kept = t_niv[j];
                if (enough) {                       // intervals are in output order: the ones that begin inside the image are a prefix
                    kept = 0;
                    while (kept < t_niv[j] && opos + slots[t_iv0[j] + kept].out < flen) kept++;
                }
                v_idx[nv] = (uint16_t)j; v_opos[nv] = (uint32_t)opos; v_ivoff[nv] = niv; v_keep[nv] = kept; nv++;
                niv += kept;
                if (enough) { complete = true; done = 1; break; }
                opos += avail; fin = t_fin[j]; last_s = s;
                if (fin) { done = 1; break; }
                s = t_next[j];
            }
            if (s >= nseg) done = 1;
            sh_s = s; sh_nv = nv; sh_done = done;
        }
        __syncthreads();
        const int nv = sh_nv;
        for (int v = warp; v < nv; v += kPlanThreads / 32) {
            const int j = v_idx[v];
            const unsigned long long o = v_opos[v];
            const uint32_t keep = v_keep[v], src = t_iv0[j], dst = (uint32_t)P.iv0 + v_ivoff[v];
            for (uint32_t i = lane; i < keep; i += 32) {
                DecIvD I = slots[src + i];
                const unsigned long long a = o + I.out;
                I.out = (uint32_t)a;
                if (a + I.len >= flen) { I.len = (uint32_t)(flen - a); I.last = 1; }
                ivs[dst + i] = I;
            }
            if (lane == 0) S[base + j].opos = (uint32_t)o;
        }
        __syncthreads();
    }
    if (tid == 0) {
        if (status == INF_OK && !complete && !fin) status = INF_SHORT;
        P.niv = (int32_t)niv;
        P.valid_len = complete ? flen : opos;
        P.done_bit = 0; P.end_bit = (complete || last_s < 0) ? 0ull : S[last_s].end_bit; P.post_err_bit = 0; P.post_err = 0; P.adler = 0;
        if (status != INF_OK) P.status = status;
    }
}

// Whole warp, once per page, behind the token that produced the last byte of the image.  zlib's inflate() — which Pillow drives one
// image row at a time — goes on from there through everything that needs no output space: end-of-block codes, block headers, empty
// stored blocks, and after a final block the Adler-32.  So a malformed header there, or a wrong checksum, fails the image in Pillow
// (if the bytes were handed to inflate() in the same call: the host checks that, api.cu decode_verdict), while anything behind the
// first literal or match is never seen.  B = bit position of the next token (in_block) or block header; last = BFINAL of the block.
__device__ void post_complete(InflMem& M, BitReader& br, uint32_t& filled, unsigned long long B, bool in_block, int last, DecPageD& P) {
    const int lane = threadIdx.x & 31;
    const unsigned long long nbits = br.n * 8ull;
    unsigned long long end_bit = 0, err_bit = 0; int err = 0;
    for (int guard = 0; guard < 64; guard++) {
        if (in_block) {
            if (B >= nbits) break;                                            // out of input: inflate() would wait for more
            topup(M, br, filled, (uint32_t)((B + 8ull * (unsigned)br.mis) >> 5));
            uint32_t w0, w1, info = 0, tok = 0;
            ring_window(M, br, B + lane, &w0, &w1);
            spec_decode<true>(M, w0, w1, &info, &tok);
            const uint32_t inf = __shfl_sync(kFull, info, 0);
            if (inf & SP_BAD) { if (B + 15 <= nbits) { err = INF_BAD_CODE; err_bit = B; } break; }
            if (!(inf & SP_EOB)) break;                                       // a literal or a match: needs output space, never decoded
            B += inf & 63u;
            if (B > nbits) break;                                             // the end-of-block code itself is cut off
            in_block = false;
            seek_bit(M, br, filled, B);
        }
        if (last) { end_bit = B; break; }
        if (B + 3 > nbits) break;
        int btype = 0, slen = 0; unsigned long long ssrc = 0;
        const int st = block_header(M, br, filled, &last, &btype, &slen, &ssrc);
        const unsigned long long after = __shfl_sync(kFull, br.bits_used(), 0);
        if (st == INF_SHORT || after > nbits) break;                          // the header is cut off
        if (st != INF_OK) { err = st; err_bit = B; break; }
        if (btype == 0) {
            if (slen > 0) break;                                              // stored bytes need output space
            B = 8ull * ssrc;
            continue;
        }
        B = after; in_block = true;
    }
    if (lane == 0) { P.end_bit = end_bit; P.post_err = err; P.post_err_bit = err_bit; }
}

// ------------------------------------------------------------------------------------------ exec: one warp per interval
__global__ void __launch_bounds__(32, 32) k_infl_exec(DecPageD* __restrict__ pages, const DecSegD* __restrict__ segs, const DecIvD* __restrict__ ivs,
                                                  int npages, int total_cap) {
    __shared__ InflMem M;
    const int lane = threadIdx.x;
    const int g = blockIdx.x;
    if (g >= total_cap) return;
    int lo = 0, hi = npages - 1;                          // page of slot g: last page with iv0 <= g
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (pages[mid].iv0 <= g) lo = mid; else hi = mid - 1; }
    DecPageD& P = pages[lo];
    const int li = g - P.iv0;
    if (P.status != 0 || li >= P.niv) return;
    const DecIvD I = ivs[g];
    BitReader br; br.init(P.z, P.zlen, M.zbuf);
    uint32_t filled = 0;
    seek_bit(M, br, filled, I.hdr_bit);
    uint16_t* __restrict__ out = P.sym + I.out;
    const unsigned long long abs0 = I.out, nbits = P.zlen * 8ull;
    const uint32_t target = I.len;
    uint32_t pos = 0;
    int status = INF_OK, last = 0;
    bool first = true;
    bool clipped = false, in_block = false;                                   // I.last: how the image ended
    unsigned long long end_B = 0;
    while (status == INF_OK && pos < target) {
        int btype = 0, slen = 0; unsigned long long ssrc = 0;
        status = block_header(M, br, filled, &last, &btype, &slen, &ssrc);
        if (status != INF_OK) break;
        unsigned long long B = __shfl_sync(kFull, br.bits_used(), 0);         // bit position of the next token
        if (first) {
            first = false;
            if (I.start_bit > B) {
                if (btype == 0) { status = INF_SEG; break; }                  // checkpoints never fall inside a stored block
                B = I.start_bit;
            }
        }
        if (btype == 0) {
            if (pos + (uint32_t)slen > target) {
                if (!I.last) { status = INF_SEG; break; }
                slen = (int)(target - pos); clipped = true;                   // the image ends inside this stored block
            }
            in_block = false; end_B = 8ull * (ssrc + (unsigned long long)slen);
            const uint8_t* s = P.z + ssrc;
            for (int k = lane; k < slen; k += 32) out[pos + k] = (uint16_t)s[k];
            pos += (uint32_t)slen;
            __syncwarp();
            if (last && pos < target) status = INF_SHORT;
            continue;
        }
        bool eob = false;
        while (status == INF_OK && !eob && pos < target) {
            // keep the input ring ahead of the parser: a batch of 32 tokens eats at most 192 bytes; the next 64 words are requested
            // now and land in the ring after the batch (latency hidden behind the parse)
            if (B > br.n * 8ull + 64ull) { status = INF_SHORT; break; }
            const uint32_t w0 = (uint32_t)((B + 8ull * (unsigned)br.mis) >> 5);
            uint32_t p0 = 0, p1 = 0; bool pf = false;
            if (filled < w0 + 96) topup(M, br, filled, w0);
            else if (filled + 64 <= w0 + kZWords) {
                pf = true;
                p0 = __ldg(br.zw + min(filled + lane, br.maxw)); p1 = __ldg(br.zw + min(filled + 32 + lane, br.maxw));
            }
            int ntok = 0;
            uint32_t room = target - pos;
            while (ntok < 32 && room > 0 && !eob && status == INF_OK) {
                uint32_t x0, x1, info = 0, tok = 0;
                ring_window(M, br, B + lane, &x0, &x1);
                spec_decode<false>(M, x0, x1, &info, &tok);
                int o = 0;
                const bool near_end = B + 2048ull > nbits;
                while (o < 32 && ntok < 32 && room > 0) {
                    uint32_t inf = __shfl_sync(kFull, info, o);
                    if ((inf & (SP_SLOW | SP_BAD | SP_EOB)) || near_end) {
                        if (inf & SP_SLOW) {
                            if (lane == o) spec_decode<true>(M, x0, x1, &info, &tok);
                            inf = __shfl_sync(kFull, info, o);
                        }
                        if (near_end && B + o + ((inf & SP_BAD) ? 48u : (inf & 63u)) > nbits) { status = INF_SHORT; break; }
                        if (inf & SP_BAD) { status = INF_BAD_CODE; break; }
                        if (inf & SP_EOB) { o += (int)(inf & 63u); eob = true; break; }
                    }
                    uint32_t sz = (inf >> 6) & 511u;
                    if (sz > room) {                                          // the probe cut at a token boundary: must land exactly ...
                        if (!I.last) { status = INF_SEG; break; }
                        if (lane == o) tok = (tok & ~511u) | room;            // ... except where the image ends inside a match
                        sz = room; clipped = true;
                    }
                    if (lane == o) M.tok[ntok] = tok;
                    ntok++; room -= sz; o += (int)(inf & 63u);
                }
                B += o;
            }
            if (status != INF_OK) break;
            __syncwarp();
            const uint32_t t = lane < ntok ? M.tok[lane] : 0u;
            const bool lit = (t >> 31) != 0u;
            const int size = lane < ntok ? (lit ? 1 : (int)(t & 511u)) : 0;
            int incl = size;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += v; }
            const int total = __shfl_sync(kFull, incl, 31);
            const uint32_t my = pos + (uint32_t)(incl - size);
            if (lit) out[my] = (uint16_t)(t & 255u);
            const bool is_match = lane < ntok && !lit;
            if (__any_sync(kFull, is_match && (unsigned long long)(t >> 9) > abs0 + my)) { status = INF_BAD_DIST; break; }
            // A match whose source lies wholly in front of this batch's output depends on nothing in the batch (on a page: the row above).
            // Those go first, four 32-symbol pieces at a time with all their loads in flight together; one at a time each piece waited
            // an L2 round trip for its own load (20 % of the kernel's stall samples).  The others follow in token order.
            const bool far = is_match && (t >> 9) >= (t & 511u) + (my - pos);
            uint32_t mf = __ballot_sync(kFull, far);
            uint32_t mm = __ballot_sync(kFull, is_match && !far);
            __syncwarp();
            {
                constexpr int U = 4;
                uint32_t fat = 0; int flen = 0, fdist = 0, fk = 0;
                while (mf || fk < flen) {
                    uint32_t ua[U]; int ul[U], ud[U], uk[U];
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        if (fk >= flen && mf) {
                            const int f = __ffs(mf) - 1; mf &= mf - 1;
                            const uint32_t tf = __shfl_sync(kFull, t, f);
                            fat = __shfl_sync(kFull, my, f); flen = (int)(tf & 511u); fdist = (int)(tf >> 9); fk = 0;
                        }
                        ua[u] = fat; ud[u] = fdist; uk[u] = fk; ul[u] = fk < flen ? flen : 0;
                        fk += 32;
                    }
                    uint16_t r[U];
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const int k = uk[u] + lane;
                        r[u] = 0;
                        if (k < ul[u]) { const int q = (int)ua[u] - ud[u] + k; r[u] = q < 0 ? (uint16_t)(256 + kWin + q) : out[q]; }
                    }
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const int k = uk[u] + lane;
                        if (k < ul[u]) out[ua[u] + k] = r[u];
                    }
                }
                __syncwarp();
            }
            while (mm) {
                const int f = __ffs(mm) - 1; mm &= mm - 1;
                const uint32_t tf = __shfl_sync(kFull, t, f);
                const uint32_t at = __shfl_sync(kFull, my, f);
                const int len = (int)(tf & 511u), dist = (int)(tf >> 9);
                if (dist == 1) {
                    // a run (most bytes of a page sit in 258-long ones): one symbol, stored four at a time.  The symbol in front of it is
                    // the previous token's if that was a literal (no load behind a store of a moment ago)
                    const uint32_t tp = __shfl_sync(kFull, t, max(f - 1, 0));
                    const uint32_t v = (f > 0 && (tp >> 31)) ? (tp & 255u) : at == 0 ? (uint32_t)(256 + kWin - 1) : (uint32_t)out[at - 1];
                    uint16_t* d = out + at;
                    const int head = min(len, (int)((4u - (uint32_t)(((uintptr_t)d >> 1) & 3u)) & 3u));
                    const int body = (len - head) >> 2, tail0 = head + 4 * body;
                    if (lane < head) d[lane] = (uint16_t)v;
                    const uint2 vv = make_uint2(v | (v << 16), v | (v << 16));
                    for (int i = lane; i < body; i += 32) reinterpret_cast<uint2*>(d + head)[i] = vv;
                    if (lane < len - tail0) d[tail0 + lane] = (uint16_t)v;
                    __syncwarp();
                    continue;
                }
                for (int k0 = 0; k0 < len; k0 += 32) {
                    const int k = k0 + lane;
                    if (k < len) {
                        // the source is the `dist` symbols in front of the match, repeated when it is shorter than the match;
                        // in front of the interval it is a symbol of its own: 256 + index into the 32 KiB window
                        const int q = (int)at - dist + (dist >= len ? k : k % dist);
                        out[at + k] = q < 0 ? (uint16_t)(256 + kWin + q) : out[q];
                    }
                }
                __syncwarp();
            }
            if (status != INF_OK) break;
            if (pf) { M.zbuf[(filled + lane) & (kZWords - 1)] = p0; M.zbuf[(filled + 32 + lane) & (kZWords - 1)] = p1; filled += 64; }
            __syncwarp();
            pos += (uint32_t)total;
        }
        if (status == INF_OK && eob && last && pos < target) status = INF_SHORT;
        if (status == INF_OK && pos < target) seek_bit(M, br, filled, B);     // the header reader continues behind the end-of-block code
        in_block = !eob; end_B = B;
    }
    if (lane == 0 && status != INF_OK) atomicMin(&P.status, status);
    if (I.last && status == INF_OK && !clipped) {
        // the image ended on a token boundary: record it and look at what inflate() would still have processed
        if (lane == 0) P.done_bit = end_B;
        post_complete(M, br, filled, end_B, in_block, last, P);
    }
}

// The last 32 KiB of every interval, in stream order: symbol -> byte through the (already concrete) 32 KiB in front of the interval.
// This is the one serial chain of a page (one step per interval), so a step is spread over a cluster of 8 CTAs (8 SMs) that meet at
// the hardware cluster barrier: 4 elements per thread, three dependent L2 round trips and one barrier per step.
constexpr int kWinCluster = 8;
__global__ void __cluster_dims__(kWinCluster, 1, 1) __launch_bounds__(1024) k_infl_window(const DecPageD* __restrict__ pages, const DecIvD* __restrict__ ivs, int n) {
    cg::cluster_group cluster = cg::this_cluster();
    const int pg = blockIdx.x / kWinCluster;
    const unsigned rank = cluster.block_rank();
    if (pg >= n) return;
    const DecPageD& P = pages[pg];
    if (P.status != 0) return;                          // the same for every CTA of the cluster
    const DecIvD* __restrict__ IV = ivs + P.iv0;
    const uint16_t* __restrict__ sym = P.sym;           // local copies: a store through P.filt could alias the descriptor otherwise
    uint8_t* filt = P.filt;
    const int niv = P.niv;
    constexpr int E = kWin / (kWinCluster * 1024);      // elements per thread per step
    unsigned long long a = niv ? IV[0].out : 0; uint32_t len = niv ? IV[0].len : 0;
    for (int s = 0; s < niv; s++) {
        const unsigned long long end = a + len;
        const unsigned long long lo = len > (uint32_t)kWin ? end - kWin : a;
        const unsigned long long p0 = lo + rank * 1024u + threadIdx.x;
        unsigned long long a_next = 0; uint32_t len_next = 0;
        if (s + 1 < niv) { a_next = IV[s + 1].out; len_next = IV[s + 1].len; }      // in flight with this step's loads
        uint32_t v[E];
#pragma unroll
        for (int i = 0; i < E; i++) { const unsigned long long p = p0 + (unsigned long long)(kWinCluster * 1024) * i; v[i] = p < end ? sym[p] : 0u; }
#pragma unroll
        for (int i = 0; i < E; i++) {         // unconditional loads (a byte symbol re-reads its own slot): nothing to branch around
            const unsigned long long p = p0 + (unsigned long long)(kWinCluster * 1024) * i;
            const bool ref = v[i] >= 256u;
            const uint32_t g = __ldcg(filt + (ref ? a - kWin + (v[i] - 256u) : (p < end ? p : lo)));
            v[i] = ref ? g : v[i];
        }
#pragma unroll
        for (int i = 0; i < E; i++) { const unsigned long long p = p0 + (unsigned long long)(kWinCluster * 1024) * i; if (p < end) filt[p] = (uint8_t)v[i]; }
        cluster.sync();                       // release / acquire at cluster scope: the next step reads these bytes through L2
        a = a_next; len = len_next;
    }
}

__global__ void __launch_bounds__(256) k_infl_resolve(const DecBatchD b) {
    const int ck = blockIdx.x;
    const DecPageD& P = b.pages[b.chunk_page[ck]];
    if (P.status != 0) return;
    const DecIvD* __restrict__ S = b.ivs + P.iv0;
    const int nseg = P.niv;
    const unsigned long long c0 = b.chunk_pos[ck], flen = P.valid_len;      // (a stream that ended early has nothing behind valid_len)
    const uint16_t* __restrict__ sym = P.sym;
    uint8_t* filt = P.filt;
    // interval of the chunk's first position (last s with out <= c0), searched once per CTA; threads walk on from there
    __shared__ int s_first;
    if (threadIdx.x == 0) {
        int lo = 0, hi = nseg - 1;
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if ((unsigned long long)S[mid].out <= c0) lo = mid; else hi = mid - 1; }
        s_first = lo;
    }
    __syncthreads();
    int s = s_first;
    unsigned long long s_end = (unsigned long long)S[s].out + S[s].len;
    for (int it = 0; it < kResolveChunk / (256 * 8); it++) {
        const unsigned long long p0 = c0 + (unsigned long long)(it * 256 + threadIdx.x) * 8ull;
        if (p0 >= flen) break;
        while (p0 >= s_end && s + 1 < nseg) { s++; s_end = (unsigned long long)S[s].out + S[s].len; }
        const int cnt = (int)min(8ull, flen - p0);
        uint16_t v8[8];
        if (cnt == 8) { const uint4 q = *reinterpret_cast<const uint4*>(sym + p0); memcpy(v8, &q, 16); }
        else for (int k = 0; k < cnt; k++) v8[k] = sym[p0 + k];
        uint8_t o8[8]; bool all = cnt == 8;
        if (cnt == 8 && p0 + 8 + kWin <= s_end) {
            // the usual group: one interval, nothing in its (already concrete) tail.  Loads are issued together; a symbol equal to
            // its left neighbour (runs are most of a page) reuses the neighbour's byte
            const uint8_t* base = filt + (unsigned long long)S[s].out - kWin - 256;
            uint32_t g[8];
#pragma unroll
            for (int k = 0; k < 8; k++) g[k] = (v8[k] >= 256u && (k == 0 || v8[k] != v8[k - 1])) ? (uint32_t)__ldg(base + v8[k]) : 0u;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const uint32_t v = v8[k];
                o8[k] = v < 256u ? (uint8_t)v : (k && v == v8[k - 1]) ? o8[k - 1] : (uint8_t)g[k];
            }
            uint2 w; memcpy(&w, o8, 8); *reinterpret_cast<uint2*>(filt + p0) = w;
            continue;
        }
        for (int k = 0; k < cnt; k++) {
            const unsigned long long p = p0 + k;
            while (p >= s_end && s + 1 < nseg) { s++; s_end = (unsigned long long)S[s].out + S[s].len; }
            const bool tail = p + kWin >= s_end;         // concrete already (k_infl_window)
            const uint32_t v = v8[k];
            uint8_t o = (uint8_t)v;
            if (!tail && v >= 256u) o = __ldg(filt + (unsigned long long)S[s].out - kWin + (v - 256u));
            if (tail) all = false;
            o8[k] = o;
            v8[k] = tail ? 1 : 0;
        }
        if (all) { uint2 w; memcpy(&w, o8, 8); *reinterpret_cast<uint2*>(filt + p0) = w; }
        else for (int k = 0; k < cnt; k++) if (!v8[k]) filt[p0 + k] = o8[k];
    }
}

// ------------------------------------------------------------------------------------------ Adler-32 of the inflated stream
// zlib checks it inside inflate() (adler32.c), so Pillow rejects a PNG whose pixels decode but whose trailer does not match.  Same
// row-partial scheme as the encoder (png_filter.cu): a CTA folds one 32 KiB chunk into (sum b, sum (n - i) b) mod 65521, a warp per
// page chains the chunks with  b += n * a + s2 ; a += s1.  A stream whose final block ended early (valid_len < filt_len: Pillow
// leaves the missing rows black) is checked over what exists, and its missing rows are written as filter type 0, zeros.
__global__ void __launch_bounds__(256) k_dec_adler_part(const DecBatchD b) {
    __shared__ uint32_t r1[8], r2[8];
    const int ck = blockIdx.x;
    const DecPageD& P = b.pages[b.chunk_page[ck]];
    if (P.status != 0) return;
    const unsigned long long c0 = b.chunk_pos[ck], vlen = P.valid_len, flen = P.filt_len;
    uint8_t* filt = P.filt;
    const int n = c0 < vlen ? (int)min((unsigned long long)kResolveChunk, vlen - c0) : 0;
    uint32_t s1 = 0; unsigned long long s2 = 0;
    const uint4* __restrict__ q = reinterpret_cast<const uint4*>(filt + c0);           // chunks start 32 KiB apart in a 256-byte aligned stream
    for (int i = threadIdx.x * 16; i < n; i += 256 * 16) {
        if (i + 16 <= n) {
            const uint4 v = q[i >> 4];
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t t = __dp4a(w[k], 0x01010101u, 0u);
                s1 += t;
                s2 += (unsigned long long)(uint32_t)(n - (i + 4 * k)) * t - __dp4a(w[k], 0x03020100u, 0u);
            }
        } else {
            for (int k = i; k < n; k++) { const uint32_t v = filt[c0 + k]; s1 += v; s2 += (unsigned long long)(uint32_t)(n - k) * v; }
        }
    }
    uint32_t s2m = (uint32_t)(s2 % 65521ull);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(kFull, s1, o); s2m += __shfl_xor_sync(kFull, s2m, o); }
    if ((threadIdx.x & 31) == 0) { r1[threadIdx.x >> 5] = s1; r2[threadIdx.x >> 5] = s2m; }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t1 = 0, t2 = 0;
        for (int w = 0; w < 8; w++) { t1 += r1[w]; t2 += r2[w]; }
        b.chunk_adler[2 * ck] = t1 % 65521u; b.chunk_adler[2 * ck + 1] = t2 % 65521u;
    }
    if (vlen < flen) {                                                                 // rows the stream never reached
        const unsigned long long rowlen = 1ull + (unsigned long long)P.w * P.c;
        const unsigned long long z0 = max(c0, vlen / rowlen * rowlen), z1 = min(flen, c0 + (unsigned long long)kResolveChunk);
        __syncthreads();
        for (unsigned long long k = z0 + threadIdx.x; k < z1; k += 256) filt[k] = 0;
    }
}

__global__ void __launch_bounds__(32) k_dec_adler_fin(const DecBatchD b) {
    DecPageD& P = b.pages[blockIdx.x];
    if (P.status != 0 || threadIdx.x != 0) return;
    const unsigned long long vlen = P.valid_len;
    uint32_t a = 1, s = 0;
    const int nck = (int)((vlen + kResolveChunk - 1) / kResolveChunk);
    for (int k = 0; k < nck; k++) {
        const uint32_t n = (uint32_t)min((unsigned long long)kResolveChunk, vlen - (unsigned long long)k * kResolveChunk);
        const uint32_t p1 = b.chunk_adler[2 * (P.chunk0 + k)], p2 = b.chunk_adler[2 * (P.chunk0 + k) + 1];
        s = (uint32_t)((s + (unsigned long long)n * a + p2) % 65521ull);
        a = (a + p1) % 65521u;
    }
    P.adler = (s << 16) | a;
}

int launch_inflate(const DecBatchD& b, cudaStream_t st) {
    if (b.npages == 0 || b.seg_total == 0) return 0;
    if (b.nscan && !b.no_scan) {
        k_infl_scan1<<<b.nscan, 256, 0, st>>>(b);
        k_infl_scan2<<<(b.surv_total + 127) / 128, 128, 0, st>>>(b);
    }
    k_infl_sort<<<b.npages, 256, 0, st>>>(b);
    int extra = 0;
    if (b.spec_total > 0) {                              // further parse units inside long blocks, then the units in order again
        k_infl_spec<<<(b.spec_total + 3) / 4, 128, 0, st>>>(b);
        k_infl_sort<<<b.npages, 256, 0, st>>>(b);
        extra = 2;
    }
    k_infl_probe<<<(b.seg_total + 3) / 4, 128, 0, st>>>(b.pages, b.segs, b.slots, b.npages, b.seg_total);
    k_infl_plan<<<b.npages, kPlanThreads, 0, st>>>(b.pages, b.segs, b.slots, b.ivs, b.npages);
    k_infl_exec<<<b.iv_total, 32, 0, st>>>(b.pages, b.segs, b.ivs, b.npages, b.iv_total);
    k_infl_window<<<b.npages * kWinCluster, 1024, 0, st>>>(b.pages, b.ivs, b.npages);
    if (b.nchunks) {
        k_infl_resolve<<<b.nchunks, 256, 0, st>>>(b);
        k_dec_adler_part<<<b.nchunks, 256, 0, st>>>(b);
    }
    k_dec_adler_fin<<<b.npages, 32, 0, st>>>(b);
    return 8 + extra + (b.nchunks ? 2 : 0);
}

// ------------------------------------------------------------------------------------------ un-filter
namespace {

constexpr int kUfInW = 36;                // words per staged input row: a TMA copy of whole 16-byte granules around the row's 32 pixels (<= 15 + 128 bytes)
constexpr int kUfOutW = 33;               // words per staged output row; odd pitch: lane r on word k of row r hits bank r + k
constexpr int kUfSin = 3, kUfSout = 2;    // stages of the input / output rings
constexpr int kUfGroup = VCP_UF_GROUP;    // consecutive bands of a page per CTA (vcp_internal.cuh)

struct __align__(16) UfSmem {
    uint32_t in[kUfSin][32][kUfInW];      // row r of a stage: the 16-byte granules that cover its 32 pixels of the chunk
    uint32_t up[kUfSin][36];              // the aligned words that cover the same 32 pixels of the last row of the band above
    uint32_t out[kUfSout][32][kUfOutW];   // row r: the chunk's pixels, byte 0 = first pixel
    uint32_t carry[32];                   // row r: the last word of the previous chunk (its tail may still be owed to global memory)
    uint64_t in_full[kUfSin], in_empty[kUfSin], out_full[kUfSout], out_empty[kUfSout];
    uint32_t ticket;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}

// aligned word `k` of the run that starts at byte address a (a itself may be misaligned); 0 outside [lo, hi)
__device__ __forceinline__ uint32_t word_at(const uint8_t* a, int k, const uint8_t* lo, const uint8_t* hi, bool l2) {
    const uint8_t* w = a - ((uintptr_t)a & 3) + 4 * k;
    if (w < lo || w + 4 > hi) return 0u;
    return l2 ? __ldcg(reinterpret_cast<const uint32_t*>(w)) : __ldg(reinterpret_cast<const uint32_t*>(w));
}

// ---- pixels as packed bytes (one register, channel k in byte k) and as 16-bit pairs (channels 0,1 / 2,3 in the halves of two registers):
//      sm_100a has VIADD.16x2, VIMNMX.16x2 and VABSDIFF4 but no packed byte compare, and 8-bit channels need 10 bits for Paeth's a + b - 2c.
template <int BPP> struct UfPx { uint32_t v[(BPP + 1) / 2]; };

template <int BPP> __device__ __forceinline__ UfPx<BPP> uf_unpack(uint32_t p) {
    UfPx<BPP> u;
    u.v[0] = __byte_perm(p, 0u, BPP == 1 ? 0x4440u : 0x4140u);
    if (BPP >= 3) u.v[(BPP + 1) / 2 - 1] = __byte_perm(p, 0u, BPP == 3 ? 0x4442u : 0x4342u);
    return u;
}
template <int BPP> __device__ __forceinline__ uint32_t uf_pack(const UfPx<BPP>& u) {
    return BPP >= 3 ? __byte_perm(u.v[0], u.v[(BPP + 1) / 2 - 1], 0x6420u) : __byte_perm(u.v[0], 0u, 0x4420u);
}
// pixel m (0..3) of the 4 * BPP dense bytes in d[]; bytes above the pixel's channels are whatever follows
template <int BPP> __device__ __forceinline__ uint32_t uf_px_get(const uint32_t* d, int m) {
    if (BPP == 1) return d[0] >> (8 * m);
    if (BPP == 2) return (m & 1) ? d[m >> 1] >> 16 : d[m >> 1];
    if (BPP == 4) return d[m];
    return m == 0 ? d[0] : m == 1 ? __funnelshift_r(d[0], d[1], 24) : m == 2 ? __funnelshift_r(d[1], d[2], 16) : d[2] >> 8;
}
// the 4 * BPP dense bytes of four packed pixels
template <int BPP> __device__ __forceinline__ void uf_px_put(const uint32_t* o, uint32_t* w) {
    if (BPP == 1) w[0] = __byte_perm(__byte_perm(o[0], o[1], 0x0040u), __byte_perm(o[2], o[3], 0x0040u), 0x5410u);
    if (BPP == 2) { w[0] = __byte_perm(o[0], o[1], 0x5410u); w[1] = __byte_perm(o[2], o[3], 0x5410u); }
    if (BPP == 3) { w[0] = __byte_perm(o[0], o[1], 0x4210u); w[1] = __byte_perm(o[1], o[2], 0x5421u); w[2] = __byte_perm(o[2], o[3], 0x6542u); }
    if (BPP == 4) { w[0] = o[0]; w[1] = o[1]; w[2] = o[2]; w[3] = o[3]; }
}
// all-ones in a half where x >= y (halves 0 .. 32767): the borrow-free difference x + 0x8000 - y has its top bit set there
__device__ __forceinline__ uint32_t uf_ge16(uint32_t x, uint32_t y) {
    uint32_t m;                                   // prmt with selector bit 3 replicates the byte's sign (__byte_perm masks that bit away)
    asm("prmt.b32 %0, %1, 0, 0xBB99;" : "=r"(m) : "r"(x + 0x80008000u - y));
    return m;
}

// one un-filter step on two channels: a = left, b = above, c = above-left, r = filtered bytes (halves 0..255).  None / Sub / Up are Paeth
// with the unused neighbours masked to 0 (Paeth(a,0,0) = a, Paeth(0,b,0) = b); Avg is selected over it by mask.  ZipDecode.c / libpng's
// rule: pa = |b-c|, pb = |a-c|, pc = |a+b-2c|; a unless pb < pa (then b), c if pc is below the smaller of the two.
template <bool HAS_AVG>
__device__ __forceinline__ uint32_t uf_step16(uint32_t a, uint32_t b, uint32_t c, uint32_t r, uint32_t ma, uint32_t mb, uint32_t mc, uint32_t mavg) {
    const uint32_t ae = a & ma, be = b & mb, ce = c & mc;
    const uint32_t pa = __vabsdiffu4(be, ce), pb = __vabsdiffu4(ae, ce);
    const uint32_t s = ae + be, c2 = ce + ce;
    const uint32_t pc = __vmaxu2(s, c2) - __vminu2(s, c2);
    const uint32_t keep_a = uf_ge16(pb, pa);
    const uint32_t ab = (ae & keep_a) | (be & ~keep_a);
    const uint32_t keep_ab = uf_ge16(pc, __vminu2(pa, pb));
    uint32_t pred = (ab & keep_ab) | (ce & ~keep_ab);
    if (HAS_AVG) pred = (pred & ~mavg) | (((a + b) >> 1) & mavg);
    return (r + pred) & 0x00FF00FFu;
}

// One band of 32 rows, three warps.  Warp 0 computes: lane = row, lane l runs l pixels behind lane l - 1, so the pixel above arrives by
// shuffle and the band advances 32 pixels (a chunk) at a time.  Warp 1 loads: per chunk one 1-D TMA copy per row into a three-stage
// ring (mbarrier complete_tx), and the row above the band from L2 once the band above has published it (acquire / release flag).
// Warp 2 stores finished chunks from a two-stage ring to global memory as aligned words and publishes the band's progress.  The warps
// meet only at mbarriers, so loading chunk j + 2, computing chunk j + 1 and storing chunk j overlap.
template <int BPP>
__device__ void unfilter_band(UfSmem& S, UfSmem* above /* the band above, if this CTA runs it */, bool below /* the band below is in this CTA */,
                              const DecPageD& P, int band, uint32_t* __restrict__ flags /* of this page */, int* bad, bool nowait) {
    const int warp = (threadIdx.x >> 5) % 3, lane = threadIdx.x & 31;
    const int W = P.w, nb = W * BPP, H = P.h;
    const int y0 = band * 32, y = y0 + lane;
    const bool row_ok = y < H;
    const uint8_t* __restrict__ F = P.filt;
    uint8_t* __restrict__ X = P.pix;
    const int nchunks = (W + 31 + 31) / 32;
    constexpr int NR = (BPP + 1) / 2;
    // row `lane`, chunk j: its 32 pixels start at a0 + 32 * BPP * j (x = 32 j - lane: lane l runs l pixels behind lane l - 1)
    const uint8_t* a0 = F + (unsigned long long)y * (nb + 1) + 1 - (long long)lane * BPP;
    const int ioff = (int)((uintptr_t)a0 & 15);                                   // the same in every chunk: a chunk is 32 * BPP bytes
    const uint32_t tma_bytes = row_ok ? (uint32_t)((ioff + 32 * BPP + 15) & ~15) : 0u;

    if (warp == 2) {
        // ------------------------------------------------------------------ stores
        // A chunk goes out a row at a time.  Inside a row whole aligned words are written: the word that straddles the start of the chunk
        // is completed from the previous chunk's last word (carry), the one that straddles its end waits for the next chunk.  Chunks
        // that touch either end of the row, and the band's last row, are written byte by byte, including the three bytes in front that
        // an earlier chunk may have left.  The last row goes first and the band's progress is published right behind it: it is all the
        // band below reads, and the release only has to wait for those few stores (lane 31 issues it: it has no word in the word path).
        S.carry[lane] = 0;
        __syncwarp();
        auto store_row = [&](int so, int j, int r) {
            uint8_t* dst = X + (unsigned long long)(y0 + r) * nb;
            const int x0 = 32 * j - r;
            const int g0 = x0 * BPP;
            if (x0 >= 3 && x0 + 32 <= W && (below || r != 31)) {
                const int mis = (int)((uintptr_t)(dst + g0) & 3);
                if (lane < 8 * BPP) {
                    const uint32_t lo = lane ? S.out[so][r][lane - 1] : S.carry[r];
                    const uint32_t v = __funnelshift_rc(lo, S.out[so][r][lane], 8 * (4 - mis));
                    *reinterpret_cast<uint32_t*>(dst + g0 - mis + 4 * lane) = v;
                }
            } else {
                const uint8_t* cb = reinterpret_cast<const uint8_t*>(&S.carry[r]) + 1;   // bytes -3 .. -1 of the chunk
                const uint8_t* ob = reinterpret_cast<const uint8_t*>(S.out[so][r]);
#pragma unroll
                for (int i = 0; i < BPP + 1; i++) {
                    const int bi = lane + 32 * i, gb = g0 - 3 + bi;
                    if (bi < 32 * BPP + 3 && gb >= 0 && gb < nb) dst[gb] = bi < 3 ? cb[bi] : ob[bi - 3];
                }
            }
        };
        const int nrows = min(32, H - y0);
        for (int j = 0; j < nchunks; j++) {
            const int so = j % kUfSout;
            mbar_wait(&S.out_full[so], (uint32_t)(j / kUfSout) & 1u);
            if (!below) {
                if (nrows == 32) store_row(so, j, 31);
                __syncwarp();
                if (lane == 31) asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(flags + band), "r"((uint32_t)(j + 1)) : "memory");
            }
#pragma unroll 4
            for (int r = 0; r < min(nrows, below ? 32 : 31); r++) store_row(so, j, r);
            __syncwarp();
            S.carry[lane] = S.out[so][lane][8 * BPP - 1];
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.out_empty[so]);
        }
        return;
    }
    if (warp == 1) {
        // ------------------------------------------------------------------ loads (up to kUfSin chunks ahead of the compute warp)
        uint32_t total_tx = tma_bytes;
#pragma unroll
        for (int d = 16; d; d >>= 1) total_tx += __shfl_xor_sync(kFull, total_tx, d);
        const uint8_t* Xend = X + (((unsigned long long)nb * H + 3ull) & ~3ull);   // the buffer is 256-byte aligned with slack behind
        uint32_t seen = 0;                                                        // lane 0: last value read from the flag of the band above
        uint32_t tail = 0;                                                        // packed: the pixel of the row above at x = 32 j
        for (int j = 0; j < nchunks; j++) {
            const int s = j % kUfSin;
            mbar_wait(&S.in_empty[s], ((uint32_t)(j / kUfSin) & 1u) ^ 1u);
            if (row_ok) bulk_g2s(&S.in[s][lane][0], a0 + (long long)32 * BPP * j - ioff, tma_bytes, &S.in_full[s]);
            if (above) {
                // the band above runs in this CTA: its last row comes out of its output ring.  Row 31 of its chunk c holds x = 32 c - 31 ..
                // 32 c; this chunk needs x = 32 j .. 32 j + 31: the last pixel of its chunk j (kept in `tail`) and the first 31 of chunk
                // j + 1.  Every chunk of the ring is read once, then handed back (out_empty counts the neighbour's storer and this warp).
                constexpr int TW = (31 * BPP) >> 2, TS = 8 * ((31 * BPP) & 3);
                if (j == 0) {
                    mbar_wait(&above->out_full[0], 0u);
                    tail = __funnelshift_r(above->out[0][31][TW], above->out[0][31][TW + 1], TS);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&above->out_empty[0]);
                }
                uint32_t v = 0, ntail = 0;
                if (j + 1 < nchunks) {
                    const int c = j + 1, so = c % kUfSout;
                    mbar_wait(&above->out_full[so], (uint32_t)(c / kUfSout) & 1u);
                    const uint32_t* ar = above->out[so][31];
                    if (lane < 8 * BPP) {
                        // up bytes 4 lane .. 4 lane + 3 = chunk bytes 4 lane - BPP ..
                        if (BPP == 4) v = lane ? ar[lane - 1] : 0u;
                        else {
                            const int o = 4 * lane - BPP;                        // >= 0 from lane 1 on
                            v = lane ? __funnelshift_r(ar[o >> 2], ar[(o >> 2) + 1], 8 * (o & 3)) : ar[0] << (8 * BPP);
                        }
                    }
                    ntail = __funnelshift_r(ar[TW], ar[TW + 1], TS);
                }
                if (lane == 0) v = BPP == 4 ? tail : (v | (tail & ((1u << (8 * (BPP & 3))) - 1u)));
                S.up[s][lane] = v;
                if (lane == 0) S.up[s][32] = 0u;
                tail = ntail;
                __syncwarp();
                if (lane == 0 && j + 1 < nchunks) mbar_arrive(&above->out_empty[(j + 1) % kUfSout]);
            } else if (band > 0) {            // the band above must have stored the pixels lane 0 will need in chunk j
                const uint32_t need = (uint32_t)min(j + 2, nchunks);
                if (lane == 0 && seen < need) {   // acquire load: what the band above stored before its release is visible after it
                    const uint32_t* f = flags + band - 1;
                    do {
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(f) : "memory");
                        if (seen >= need || nowait) break;
                        __nanosleep(32);
                    } while (true);
                }
                __syncwarp();
                const uint8_t* u = X + (unsigned long long)(y0 - 1) * nb + (long long)32 * j * BPP;
                S.up[s][lane] = word_at(u, lane, X, Xend, true);
                if (lane == 0) S.up[s][32] = BPP == 4 ? word_at(u, 32, X, Xend, true) : 0u;
            }
            __syncwarp();
            if (lane == 0) mbar_expect_tx(&S.in_full[s], total_tx);                 // the one arrival of this phase + the bytes in flight
        }
        return;
    }

    // ---------------------------------------------------------------------- compute
    int ft = row_ok ? F[(unsigned long long)y * (nb + 1)] : 0;
    if (ft > 4) { *bad = 1; ft = 0; }
    // filter type as masks: which neighbours the Paeth core sees (none: None; a: Sub; b: Up; all: Paeth), and Avg
    const uint32_t ma = (ft == 1 || ft == 4) ? ~0u : 0u, mb = (ft == 2 || ft == 4) ? ~0u : 0u, mc = ft == 4 ? ~0u : 0u, mavg = ft == 3 ? ~0u : 0u;
    const bool simple = __all_sync(kFull, ft <= 2);
    const bool has_avg = __any_sync(kFull, ft == 3);
    const int in_w = ioff >> 2, in_sh = 8 * (ioff & 3);
    uint32_t curp = 0;                        // packed: my pixel x - 1
    UfPx<BPP> au, cu;                         // 16-bit pairs: my pixel x - 1, the pixel above x - 1
#pragma unroll
    for (int q = 0; q < NR; q++) au.v[q] = cu.v[q] = 0;
    for (int j = 0; j < nchunks; j++) {
        const int s = j % kUfSin, so = j % kUfSout;
        mbar_wait(&S.in_full[s], (uint32_t)(j / kUfSin) & 1u);
        mbar_wait(&S.out_empty[so], ((uint32_t)(j / kUfSout) & 1u) ^ 1u);
        // lane s holds pixel s of the row above (packed), handed to lane 0 by a broadcast at step s
        uint32_t upv = 0;
        if (band > 0) {
            const int ub = (above ? 0 : (int)((uintptr_t)(X + (unsigned long long)(y0 - 1) * nb + (long long)32 * j * BPP) & 3)) + lane * BPP;
            upv = __funnelshift_r(S.up[s][ub >> 2], S.up[s][(ub >> 2) + 1], 8 * (ub & 3));
        }
        const uint32_t* inw = S.in[s][lane] + in_w;
        uint32_t* outw = S.out[so][lane];
        const bool ramp = j == 0;             // only in the first chunk a lane may not have started (x < 0): left and upper-left of pixel 0 are 0
        uint32_t wprev = inw[0];
        auto body = [&](auto simple_c, auto avg_c, auto ramp_c) {
#pragma unroll 1
            for (int i = 0; i < 8; i++) {
                uint32_t d[BPP], o[4];
#pragma unroll
                for (int k = 0; k < BPP; k++) { const uint32_t w = inw[BPP * i + k + 1]; d[k] = __funnelshift_r(wprev, w, in_sh); wprev = w; }
#pragma unroll
                for (int m = 0; m < 4; m++) {
                    const int sidx = 4 * i + m;
                    const uint32_t bs = __shfl_up_sync(kFull, curp, 1), b0 = __shfl_sync(kFull, upv, sidx);
                    const uint32_t bp = lane == 0 ? b0 : bs;
                    const uint32_t rp = uf_px_get<BPP>(d, m);
                    if (decltype(simple_c)::value) {
                        // no Avg / Paeth row in the band (blank paper is all Up): the predictor is a or b, four channels at a time
                        o[m] = __vadd4(rp, (curp & ma) | (bp & mb));
                    } else {
                        const UfPx<BPP> bu = uf_unpack<BPP>(bp), ru = uf_unpack<BPP>(rp);
                        UfPx<BPP> ou;
#pragma unroll
                        for (int q = 0; q < NR; q++) ou.v[q] = uf_step16<decltype(avg_c)::value>(au.v[q], bu.v[q], cu.v[q], ru.v[q], ma, mb, mc, mavg);
                        o[m] = uf_pack<BPP>(ou);
                        au = ou; cu = bu;
                    }
                    curp = o[m];
                    if (decltype(ramp_c)::value && sidx - lane < 0) {
                        curp = 0;
#pragma unroll
                        for (int q = 0; q < NR; q++) au.v[q] = cu.v[q] = 0;
                    }
                }
                uint32_t w[BPP];
                uf_px_put<BPP>(o, w);
#pragma unroll
                for (int k = 0; k < BPP; k++) outw[BPP * i + k] = w[k];
            }
        };
        if (ramp) { if (simple) body(std::true_type{}, std::false_type{}, std::true_type{}); else body(std::false_type{}, std::true_type{}, std::true_type{}); }
        else if (simple) body(std::true_type{}, std::false_type{}, std::false_type{});
        else if (has_avg) body(std::false_type{}, std::true_type{}, std::false_type{});
        else body(std::false_type{}, std::false_type{}, std::false_type{});
        __syncwarp();
        if (lane == 0) { mbar_arrive(&S.in_empty[s]); mbar_arrive(&S.out_full[so]); }
    }
}

}  // namespace

__global__ void __launch_bounds__(96 * kUfGroup) k_unfilter(const DecBatchD b) {
    extern __shared__ __align__(16) unsigned char uf_smem[];
    UfSmem* S = reinterpret_cast<UfSmem*>(uf_smem);
    // A CTA runs kUfGroup consecutive bands of one page (a finished group frees its slot at once): inside the group the last row of a
    // band reaches the band below through shared memory, between groups through global memory and a flag.  Groups are handed out by
    // ticket, group-major over the pages of the batch (group 0 of every page, then group 1, ...): a band only waits on one that holds an
    // earlier ticket, and the resident CTAs are the pipeline fronts of all pages rather than all bands of a few.
    const int slot = threadIdx.x / 96;
    if (threadIdx.x == 0) S[0].ticket = atomicAdd(b.counters, 1u);
    if (threadIdx.x % 96 == 0) {
        for (int i = 0; i < kUfSin; i++) { mbar_init(&S[slot].in_full[i], 1); mbar_init(&S[slot].in_empty[i], 1); }
        for (int i = 0; i < kUfSout; i++) mbar_init(&S[slot].out_full[i], 1);
    }
    __syncthreads();
    const uint32_t t = S[0].ticket;
    if ((int)t >= b.ngroups) return;
    DecPageD& P = b.pages[b.band_page[t]];
    const int band = (int)b.band_idx[t] + slot, pbands = (P.h + 31) / 32;
    uint32_t* flags = b.band_flag + P.band0;
    const bool live = band < pbands, below = slot + 1 < kUfGroup && band + 1 < pbands;
    if (threadIdx.x % 96 == 0)                            // a ring stage is free again when the storer and, if there is one, the neighbour below have read it
        for (int i = 0; i < kUfSout; i++) mbar_init(&S[slot].out_empty[i], below ? 2 : 1);
    __syncthreads();
    if (!live) return;
    int bad = 0;
    if (P.status != 0) {                                  // a skipped band still releases the bands waiting on it
        if (threadIdx.x % 96 == 0) *(volatile uint32_t*)(flags + band) = 0xffffffffu;
        return;
    }
    UfSmem* above = slot > 0 ? &S[slot - 1] : nullptr;
    switch (P.c) {
        case 1: unfilter_band<1>(S[slot], above, below, P, band, flags, &bad, b.dbg_nowait != 0); break;
        case 2: unfilter_band<2>(S[slot], above, below, P, band, flags, &bad, b.dbg_nowait != 0); break;
        case 3: unfilter_band<3>(S[slot], above, below, P, band, flags, &bad, b.dbg_nowait != 0); break;
        default: unfilter_band<4>(S[slot], above, below, P, band, flags, &bad, b.dbg_nowait != 0); break;
    }
    if (threadIdx.x % 96 < 32 && __any_sync(kFull, bad) && threadIdx.x % 96 == 0) atomicMin(&P.status, (int)INF_BAD_FILTER);
}

int launch_unfilter(const DecBatchD& b, cudaStream_t st) {
    if (b.ngroups == 0) return 0;
    static std::atomic<int> attr_by_dev[kMaxDevices];
    int dev = 0;
    cudaGetDevice(&dev);
    const int smem = (int)sizeof(UfSmem) * kUfGroup;
    if (dev >= 0 && dev < kMaxDevices && !attr_by_dev[dev].load(std::memory_order_acquire)) {
        cudaFuncSetAttribute(k_unfilter, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attr_by_dev[dev].store(1, std::memory_order_release);
    }
    k_unfilter<<<b.ngroups, 96 * kUfGroup, smem, st>>>(b);
    return 1;
}

}  // namespace vcp
