// png_decode.cu — zlib inflate and PNG un-filtering on the GPU (the reverse of deflate_*.cu / png_filter.cu).
//
// Replaces what Pillow runs for Image.open(png).load(): zlib inflate() driven by libImaging/ZipDecode.c and its per-row
// un-filter (None/Sub/Up/Avg/Paeth).  SURVEY.md §8 f-4: re-reading the reference's images/page_###.png cache
// (backend/README.md:238-243) and validating this library's own PNGs at speed.  Restated for tests by
// oracle/restate.py:png_unfilter and Python's zlib.
//
//   k_inflate   one warp per page.  Lane 0 parses the bit stream (any block type: stored, fixed, dynamic; multi-block),
//               decoding through a 10-bit lookup table in shared memory with a canonical bit-by-bit path for longer codes;
//               every token is broadcast and executed by the whole warp (matches are copied 32 bytes per step, periodic
//               for distances < 32).  Compressed pages have few tokens per byte, so the serial parse is short; the
//               parallelism of a batch is its page count.
//   k_unfilter  one warp per page, rows in order.  None / Up are element-wise; Sub is a per-channel prefix sum (segment sums +
//               warp scan); Avg and Paeth are true recurrences and run one lane per channel.
// Every loop is bounded by the stream / output length: malformed input ends in a status, never in a hang.
#include "vcp_internal.cuh"

namespace vcp {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kFastBits = 10;
constexpr int kInflWarps = 2;

enum { INF_OK = 0, INF_BAD_HEADER = -101, INF_BAD_BLOCK = -102, INF_BAD_CODE = -103, INF_OVERRUN = -104, INF_BAD_DIST = -105,
       INF_SHORT = -106, INF_BAD_FILTER = -107 };

struct InflMem {
    uint16_t fast_ll[1 << kFastBits];     // (symbol << 4) | length, 0 = not in the fast table
    uint16_t fast_d[1 << kFastBits];
    uint16_t sorted_ll[288], sorted_d[32];
    uint16_t cnt_ll[16], cnt_d[16];
    uint8_t lens[320];
    uint32_t tok[32];                     // one batch of tokens: literal = bit 31 | byte, match = distance << 9 | length
};

// LSB-first bit reader over a byte stream that has >= 8 addressable bytes of slack behind it; refills 32 bits at a time from two
// aligned words (one of them already in a register from the previous refill)
struct BitReader {
    const uint32_t* zw; int mis;                 // aligned word pointer of the stream start, byte misalignment 0..3
    unsigned long long n, pos;                   // stream length, bytes fetched so far
    unsigned long long buf; int cnt; bool over;
    uint32_t nextw;                              // aligned word (pos + mis) / 4, prefetched
    __device__ void init(const uint8_t* z, unsigned long long len) {
        mis = (int)((uintptr_t)z & 3); zw = reinterpret_cast<const uint32_t*>(z - mis);
        n = len; pos = 0; buf = 0; cnt = 0; over = false;
        nextw = __ldg(zw);
    }
    __device__ void refill() {
        if (cnt <= 32) {
            const unsigned long long w = (pos + mis) >> 2;
            const uint32_t hi = __ldg(zw + w + 1);
            const uint32_t v = __funnelshift_r(nextw, hi, 8 * (int)((pos + mis) & 3));   // bytes [pos, pos + 4)
            nextw = hi;
            buf |= (unsigned long long)v << cnt; cnt += 32; pos += 4;
            if (pos > n + 12) over = true;
        }
    }
    __device__ uint32_t peek(int k) { return (uint32_t)(buf & ((1ull << k) - 1ull)); }
    __device__ void drop(int k) { buf >>= k; cnt -= k; }
    __device__ uint32_t bits(int k) { if (cnt < k) refill(); const uint32_t v = peek(k); drop(k); return v; }
    __device__ unsigned long long byte_pos() const { return pos - (unsigned long long)(cnt >> 3); }   // first byte not consumed (cnt % 8 == 0)
    __device__ void seek(unsigned long long p) { pos = p; buf = 0; cnt = 0; nextw = __ldg(zw + ((p + mis) >> 2)); }
};

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

}  // namespace

__global__ void __launch_bounds__(kInflWarps * 32) k_inflate(DecPageD* __restrict__ pages, int n) {
    __shared__ InflMem mem[kInflWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pg = blockIdx.x * kInflWarps + warp;
    if (pg >= n) return;
    DecPageD& P = pages[pg];
    if (P.status != 0) return;
    InflMem& M = mem[warp];
    uint8_t* __restrict__ out = P.filt;
    const unsigned long long cap = P.filt_len;
    unsigned long long pos = 0;
    BitReader br; br.init(P.z, P.zlen);
    int status = INF_OK;
    if (lane == 0) {
        const uint32_t cmf = br.bits(8), flg = br.bits(8);
        if ((cmf & 15) != 8 || ((cmf << 8) | flg) % 31 != 0 || (flg & 32)) status = INF_BAD_HEADER;
    }
    status = __shfl_sync(kFull, status, 0);
    int last = 0;
    while (status == INF_OK && !last) {
        int btype = 0;
        if (lane == 0) { last = (int)br.bits(1); btype = (int)br.bits(2); }
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
            for (int k = lane; k < len; k += 32) out[pos + k] = P.z[src + k];
            pos += len;
            __syncwarp();
            continue;
        }
        if (btype == 3) { status = INF_BAD_BLOCK; break; }
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
        // ---- tokens: lane 0 decodes a batch of up to 32 into shared memory, the warp executes it: sizes are scanned, all
        //      literals are stored at once, matches run in order (each copied by the whole warp)
        bool eob = false;
        while (status == INF_OK && !eob) {
            int ntok = 0;
            if (lane == 0) {
                for (; ntok < 32; ntok++) {
                    const int s = decode_sym(br, M.fast_ll, M.cnt_ll, M.sorted_ll);
                    if (s < 0 || br.over) { status = INF_BAD_CODE; break; }
                    if (s < 256) { M.tok[ntok] = 0x80000000u | (uint32_t)s; continue; }
                    if (s == 256) { eob = true; break; }
                    if (s > 285) { status = INF_BAD_CODE; break; }
                    const int ls = s - 257;
                    const int len = kLenBase[ls] + (int)br.bits(kLenExtra[ls]);
                    const int ds = decode_sym(br, M.fast_d, M.cnt_d, M.sorted_d);
                    if (ds < 0 || ds > 29) { status = INF_BAD_CODE; break; }
                    const int dist = kDistBase[ds] + (int)br.bits(kDistExtra[ds]);
                    M.tok[ntok] = ((uint32_t)dist << 9) | (uint32_t)len;
                }
            }
            status = __shfl_sync(kFull, status, 0); ntok = __shfl_sync(kFull, ntok, 0); eob = __shfl_sync(kFull, (int)eob, 0) != 0;
            if (status != INF_OK) break;
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
            if (lit && lane < ntok) out[my] = (uint8_t)t;
            uint32_t mm = __ballot_sync(kFull, lane < ntok && !lit);
            __syncwarp();
            while (mm) {
                const int f = __ffs(mm) - 1; mm &= mm - 1;
                const uint32_t tf = __shfl_sync(kFull, t, f);
                const unsigned long long at = __shfl_sync(kFull, my, f);
                const int len = (int)(tf & 511u), dist = (int)(tf >> 9);
                if ((unsigned long long)dist > at) { status = INF_BAD_DIST; break; }
                if (dist >= len) {                                  // no overlap: every byte's source already exists
                    for (int k = lane; k < len; k += 32) out[at + k] = out[at + k - dist];
                } else if (dist >= 32) {
                    for (int k0 = 0; k0 < len; k0 += 32) {
                        const int k = k0 + lane;
                        if (k < len) out[at + k] = out[at + k - dist];
                        __syncwarp();
                    }
                } else {                                            // periodic: the source is the `dist` bytes in front
                    for (int k = lane; k < len; k += 32) out[at + k] = out[at - dist + (k % dist)];
                }
                __syncwarp();
            }
            if (status != INF_OK) break;
            pos += total;
        }
    }
    if (status == INF_OK && pos != cap) status = INF_SHORT;
    if (lane == 0) P.status = status;
}

int launch_inflate(DecPageD* d_pages, int n, cudaStream_t st) {
    if (n == 0) return 0;
    k_inflate<<<(n + kInflWarps - 1) / kInflWarps, kInflWarps * 32, 0, st>>>(d_pages, n);
    return 1;
}

// ------------------------------------------------------------------------------------------ un-filter
__global__ void __launch_bounds__(kInflWarps * 32) k_unfilter(DecPageD* __restrict__ pages, int n) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pg = blockIdx.x * kInflWarps + warp;
    if (pg >= n) return;
    DecPageD& P = pages[pg];
    if (P.status != 0) return;
    const int bpp = P.c, nb = P.w * bpp;
    const uint8_t* __restrict__ F = P.filt;
    uint8_t* __restrict__ X = P.pix;
    int status = 0;
    for (int y = 0; y < P.h; y++) {
        const uint8_t* r = F + (long long)y * (nb + 1) + 1;
        const int ft = F[(long long)y * (nb + 1)];
        uint8_t* x = X + (long long)y * nb;
        const uint8_t* up = y ? x - nb : nullptr;
        if (ft == 0) {
            for (int i = lane; i < nb; i += 32) x[i] = r[i];
        } else if (ft == 2) {
            for (int i = lane; i < nb; i += 32) x[i] = (uint8_t)(r[i] + (up ? up[i] : 0));
        } else if (ft == 1) {
            // per-channel prefix sums: lane l owns pixels [l*S, (l+1)*S)
            const int S = (P.w + 31) / 32, p0 = min(P.w, lane * S), p1 = min(P.w, p0 + S);
            uint32_t acc[4] = {0, 0, 0, 0};
            for (int px = p0; px < p1; px++)
                for (int ch = 0; ch < bpp; ch++) acc[ch] += r[px * bpp + ch];
            uint32_t pre[4];
            for (int ch = 0; ch < 4; ch++) {
                uint32_t v = acc[ch];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(kFull, v, o); if (lane >= o) v += t; }
                pre[ch] = v - acc[ch];
            }
            for (int px = p0; px < p1; px++)
                for (int ch = 0; ch < bpp; ch++) { pre[ch] += r[px * bpp + ch]; x[px * bpp + ch] = (uint8_t)pre[ch]; }
        } else if (ft == 3 || ft == 4) {
            if (lane < bpp) {
                int a = 0, c = 0;                 // left, upper-left of this channel
                for (int i = lane; i < nb; i += bpp) {
                    const int b = up ? up[i] : 0;
                    int pred;
                    if (ft == 3) pred = (a + b) >> 1;
                    else {
                        const int pp = a + b - c, pa = abs(pp - a), pb = abs(pp - b), pc = abs(pp - c);
                        pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                    }
                    const int v = (r[i] + pred) & 255;
                    x[i] = (uint8_t)v;
                    a = v; c = b;
                }
            }
        } else {
            status = INF_BAD_FILTER;
            break;
        }
        __syncwarp();                             // the next row reads this one
    }
    if (lane == 0 && status) P.status = status;
}

int launch_unfilter(DecPageD* d_pages, int n, cudaStream_t st) {
    if (n == 0) return 0;
    k_unfilter<<<(n + kInflWarps - 1) / kInflWarps, kInflWarps * 32, 0, st>>>(d_pages, n);
    return 1;
}

}  // namespace vcp
