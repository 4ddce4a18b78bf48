// deflate_huff.cu — dynamic-Huffman coding of the LZ77 token stream and zlib stream layout.
//
// Replaces zlib's trees.c (build_tree / gen_bitlen / scan_tree / send_all_trees / compress_block / bi_flush)
// and the stream framing of deflate.c as driven by Pillow's ZipEncode.c under `page_image.save(...)`
// (backend/app/pipeline/pdf_extract.py:130).  Like the LZ stage the bytes are not zlib's bytes; the contract
// is a valid zlib stream that inflates to the exact filtered rows and is <= 1.05 x Pillow's default size.
// tests/model/deflate_model.c (dm_huff_lengths, dm_canonical, dm_huff_block) states the same arithmetic
// sequentially; tests compare byte-for-byte with it and inflate with zlib.
//
// One deflate block = 512 KiB of filtered stream = one IDAT chunk, ended with an empty stored block so that
// blocks are byte aligned and independent.  Four kernels:
//   k_huff_build    CTA per block: sum the sub-chunk histograms, rank-sort, two-queue Huffman merge,
//                   length limiting (Kraft repair), canonical codes, code-length RLE + its 7-bit code, header
//                   bits; exact bit size of every sub-chunk from its histogram -> bit offsets, payload bytes,
//                   stored-block fallback decision.
//   k_layout        one CTA: scan payload sizes -> chunk positions, per-page PNG and base64 extents.
//   k_payload_init  CTA per block: zero the payload, write zlib header, block header bits, EOB, sync marker /
//                   Adler-32; stored blocks are copied here.
//   k_huff_emit     CTA per sub-chunk: tokens -> codes, block-wide scan of bit lengths, bits OR-ed into a
//                   shared staging tile, whole words stored coalesced (atomicOr only on the two boundary words).
#include "vcp_internal.cuh"

namespace vcp {

namespace {

constexpr int kBuildThreads = 160;     // 12 CTAs per SM: every block of a 64-page batch is resident at once
constexpr int kEmitThreads = 256;

__device__ __forceinline__ int len_sym(int len) {      // 3..258 -> 0..28
    const int v = len - 3;
    if (v < 8) return v;
    if (len == 258) return 28;
    const int n = 31 - __clz(v);
    return 4 * (n - 1) + ((v >> (n - 2)) & 3);
}
__device__ __forceinline__ int len_extra(int ls) { return (ls < 8 || ls == 28) ? 0 : (ls >> 2) - 1; }
__device__ __forceinline__ int dist_sym(int dist) {    // 1..32768 -> 0..29
    const int v = dist - 1;
    if (v < 4) return v;
    const int n = 31 - __clz(v);
    return 2 * n + ((v >> (n - 1)) & 1);
}
__device__ __forceinline__ int dist_extra(int ds) { return ds < 4 ? 0 : (ds >> 1) - 1; }

__device__ __forceinline__ uint32_t bitrev(uint32_t v, int n) { return __brev(v) >> (32 - n); }

// ---- serial part of the length-limited Huffman construction (one thread).
// order[0..m): symbols with freq > 0 sorted by (freq, symbol); freq[] patched so that m >= 2.
// w/parent: 2*m scratch entries.  Writes lens[sym] for the m used symbols (others stay 0).
__device__ void tree_serial(const uint32_t* freq, const uint16_t* order, int m, int maxbits,
                            uint8_t* lens, uint32_t* w, uint16_t* parent) {
    for (int i = 0; i < m; i++) w[i] = freq[order[i]];
    int li = 0, ii = m, nn = m;
    while (nn < 2 * m - 1) {
        int pick0, pick1;
        if (li < m && (ii >= nn || w[li] <= w[ii])) pick0 = li++; else pick0 = ii++;
        if (li < m && (ii >= nn || w[li] <= w[ii])) pick1 = li++; else pick1 = ii++;
        w[nn] = w[pick0] + w[pick1];
        parent[pick0] = (uint16_t)nn; parent[pick1] = (uint16_t)nn;
        nn++;
    }
    // depths: reuse w[] as depth storage from the root down (w of a node is dead once its parent is formed)
    int cnt[16];
#pragma unroll
    for (int l = 0; l < 16; l++) cnt[l] = 0;
    w[2 * m - 2] = 0;
    for (int i = 2 * m - 3; i >= 0; i--) {
        const uint32_t d = w[parent[i]] + 1;
        w[i] = d;
        if (i < m) cnt[d > (uint32_t)maxbits ? maxbits : (int)d]++;
    }
    long long K = 0;
    for (int l = 1; l <= maxbits; l++) K += (long long)cnt[l] << (maxbits - l);
    long long excess = K - (1ll << maxbits);
    while (excess > 0) {
        int l = maxbits - 1;
        while (cnt[l] == 0) l--;
        cnt[l]--; cnt[l + 1]++;
        excess -= 1ll << (maxbits - l - 1);
    }
    while (excess < 0) {
        bool done = false;
        for (int l = maxbits; l >= 2; l--) {
            const long long cost = 1ll << (maxbits - l);
            if (cnt[l] > 0 && cost <= -excess) { cnt[l]--; cnt[l - 1]++; excess += cost; done = true; break; }
        }
        if (!done) break;
    }
    int idx = m - 1;
    for (int l = 1; l <= maxbits; l++)
        for (int c = 0; c < cnt[l]; c++) lens[order[idx--]] = (uint8_t)l;
}

// serial canonical code assignment (bit-reversed), n small or called by one thread
__device__ void canonical_serial(const uint8_t* lens, int n, int maxbits, uint16_t* codes) {
    int cnt[17]; uint32_t next[17];
    for (int l = 0; l <= 16; l++) cnt[l] = 0;
    for (int i = 0; i < n; i++) cnt[lens[i]]++;
    cnt[0] = 0;
    uint32_t c = 0;
    for (int l = 1; l <= maxbits; l++) { c = (c + cnt[l - 1]) << 1; next[l] = c; }
    for (int i = 0; i < n; i++) {
        const int l = lens[i];
        codes[i] = l ? (uint16_t)bitrev(next[l]++, l) : (uint16_t)0;
    }
}

struct BitW {
    uint8_t* out; uint64_t acc; int n; uint32_t total;
    __device__ void put(uint32_t v, int nb) {
        acc |= (uint64_t)v << n; n += nb; total += nb;
        while (n >= 8) { *out++ = (uint8_t)acc; acc >>= 8; n -= 8; }
    }
    __device__ void flush() { if (n > 0) { *out++ = (uint8_t)acc; acc = 0; n = 0; } }
};

__device__ int rle_lengths(const uint8_t* L, int n, uint16_t* out) {
    int k = 0, i = 0;
    while (i < n) {
        const int v = L[i]; int r = 1;
        while (i + r < n && L[i + r] == v) r++;
        i += r;
        if (v == 0) {
            while (r >= 11) { const int t = r > 138 ? 138 : r; out[k++] = (uint16_t)(18 | ((t - 11) << 8)); r -= t; }
            if (r >= 3) { out[k++] = (uint16_t)(17 | ((r - 3) << 8)); r = 0; }
            while (r-- > 0) out[k++] = 0;
        } else {
            out[k++] = (uint16_t)v; r--;
            while (r >= 3) { const int t = r > 6 ? 6 : r; out[k++] = (uint16_t)(16 | ((t - 3) << 8)); r -= t; }
            while (r-- > 0) out[k++] = (uint16_t)v;
        }
    }
    return k;
}

__constant__ uint8_t kClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

}  // namespace

// ------------------------------------------------------------------------------------------ build
namespace {

// two-queue Huffman merge over m sorted leaves (one thread): fills parent[0 .. 2m-2); queue heads live in registers
__device__ void merge_serial(const uint32_t* freq, const uint16_t* order, int m, uint32_t* w, uint16_t* parent) {
    for (int i = 0; i < m; i++) w[i] = freq[order[i]];
    constexpr uint32_t INF = 0xFFFFFFFFu;
    int li = 0, ii = m, nn = m;
    uint32_t wl = w[0], wi = INF;
    while (nn < 2 * m - 1) {
        int p0, p1; uint32_t v0, v1;
        if (li < m && wl <= wi) { p0 = li++; v0 = wl; wl = li < m ? w[li] : INF; } else { p0 = ii++; v0 = wi; wi = ii < nn ? w[ii] : INF; }
        if (li < m && wl <= wi) { p1 = li++; v1 = wl; wl = li < m ? w[li] : INF; } else { p1 = ii++; v1 = wi; wi = ii < nn ? w[ii] : INF; }
        const uint32_t sum = v0 + v1;
        w[nn] = sum; parent[p0] = (uint16_t)nn; parent[p1] = (uint16_t)nn;
        if (ii == nn) wi = sum;                       // the internal queue was empty: the new node is its head
        nn++;
    }
    parent[2 * m - 2] = (uint16_t)(2 * m - 2);        // root points at itself
}

// clamp the per-length leaf counts to maxbits (Kraft repair), as dm_huff_lengths does (one thread)
__device__ void kraft_repair(int* cnt, int maxbits) {
    long long K = 0;
    for (int l = 1; l <= maxbits; l++) K += (long long)cnt[l] << (maxbits - l);
    long long excess = K - (1ll << maxbits);
    while (excess > 0) {
        int l = maxbits - 1;
        while (cnt[l] == 0) l--;
        cnt[l]--; cnt[l + 1]++;
        excess -= 1ll << (maxbits - l - 1);
    }
    while (excess < 0) {
        bool done = false;
        for (int l = maxbits; l >= 2; l--) {
            const long long cost = 1ll << (maxbits - l);
            if (cnt[l] > 0 && cost <= -excess) { cnt[l]--; cnt[l - 1]++; excess += cost; done = true; break; }
        }
        if (!done) break;
    }
}

}  // namespace

__global__ void __launch_bounds__(kBuildThreads, 10) k_huff_build(BatchD B) {   // <= 40 registers: all 1408 blocks of a 64-page batch resident at once
    __shared__ uint32_t freq[kCodeStride];          // [0,286) lit/len, [286,316) dist (patched copies)
    __shared__ uint16_t order_ll[kNumLL], order_d[kNumD];
    __shared__ uint8_t lens[kCodeStride];
    __shared__ uint16_t codes[kCodeStride];
    __shared__ uint32_t w_ll[2 * kNumLL]; __shared__ uint16_t par_ll[2 * kNumLL]; __shared__ uint16_t dep_ll[2 * kNumLL];
    __shared__ uint32_t w_d[2 * kNumD];   __shared__ uint16_t par_d[2 * kNumD];   __shared__ uint16_t dep_d[2 * kNumD];
    __shared__ int m_ll, m_d;
    __shared__ int cnt_ll[16], cnt_d[16];
    __shared__ uint32_t cost[kCodeStride];          // bits per occurrence of each symbol (code + extra)
    __shared__ unsigned long long sub_bits[kBlockBytes / kSubBytes];
    __shared__ uint32_t hdr_bits_s;

    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const BlockD blk = B.blocks[b];

    // ---- block histogram
    for (int t = tid; t < kCodeStride; t += kBuildThreads) {
        uint32_t s = 0;
        if (t < kHistSize) {
            for (int j = 0; j < blk.nsub; j++) s += B.sub_hist[(size_t)(blk.sub0 + j) * kHistSize + t];
            if (t == 256) s += 1;                    // EOB
        }
        freq[t] = s; lens[t] = 0; codes[t] = 0; cost[t] = 0;
    }
    if (tid < 16) { cnt_ll[tid] = 0; cnt_d[tid] = 0; }
    __syncthreads();
    // ---- at least two used symbols per alphabet (dm_huff_lengths); symbol counts
    if (tid == 0) {
        int used = 0; for (int i = 0; i < kNumLL; i++) used += freq[i] != 0;   // EOB makes used >= 1
        if (used == 1) { if (freq[0]) freq[1] = 1; else freq[0] = 1; used = 2; }
        m_ll = used;
    } else if (tid == 32) {
        uint32_t* f = freq + kNumLL;
        int used = 0; for (int i = 0; i < kNumD; i++) used += f[i] != 0;
        if (used == 0) { f[0] = 1; f[1] = 1; used = 2; }
        else if (used == 1) { if (f[0]) f[1] = 1; else f[0] = 1; used = 2; }
        m_d = used;
    }
    __syncthreads();
    // ---- rank sort by (freq, symbol) among used symbols
    for (int t = tid; t < kHistSize; t += kBuildThreads) {
        const bool isd = t >= kNumLL;
        const int lo = isd ? kNumLL : 0, hi = isd ? kHistSize : kNumLL;
        const uint32_t f = freq[t];
        if (f) {
            int r = 0;
            for (int j = lo; j < hi; j++) { const uint32_t g = freq[j]; r += (g != 0) && (g < f || (g == f && j < t)); }
            if (isd) order_d[r] = (uint16_t)(t - kNumLL); else order_ll[r] = (uint16_t)t;
        }
    }
    __syncthreads();
    // ---- the two merges, one thread each (different warps)
    if (tid == 0) merge_serial(freq, order_ll, m_ll, w_ll, par_ll);
    else if (tid == 32) merge_serial(freq + kNumLL, order_d, m_d, w_d, par_d);
    __syncthreads();
    // ---- leaf depths by pointer jumping (9 rounds cover depth < 512), both trees at once
    {
        const int n1 = 2 * m_ll - 1, n2 = 2 * m_d - 1, nt = n1 + n2;
        for (int t = tid; t < nt; t += kBuildThreads) {
            if (t < n1) dep_ll[t] = (t == n1 - 1) ? 0 : 1; else dep_d[t - n1] = (t - n1 == n2 - 1) ? 0 : 1;
        }
        __syncthreads();
        constexpr int kPer = (2 * kNumLL + 2 * kNumD + kBuildThreads - 1) / kBuildThreads;
        for (int round = 0; round < 9; round++) {
            uint16_t nd[kPer], np[kPer];
#pragma unroll
            for (int k = 0; k < kPer; k++) {
                const int t = tid + k * kBuildThreads;
                if (t < nt) {
                    uint16_t* dep = t < n1 ? dep_ll : dep_d; uint16_t* par = t < n1 ? par_ll : par_d;
                    const int i = t < n1 ? t : t - n1;
                    const int p = par[i];
                    nd[k] = (uint16_t)(dep[i] + dep[p]); np[k] = par[p];
                }
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < kPer; k++) {
                const int t = tid + k * kBuildThreads;
                if (t < nt) {
                    uint16_t* dep = t < n1 ? dep_ll : dep_d; uint16_t* par = t < n1 ? par_ll : par_d;
                    const int i = t < n1 ? t : t - n1;
                    dep[i] = nd[k]; par[i] = np[k];
                }
            }
            __syncthreads();
        }
        for (int t = tid; t < m_ll + m_d; t += kBuildThreads) {
            if (t < m_ll) atomicAdd(&cnt_ll[min((int)dep_ll[t], 15)], 1);
            else atomicAdd(&cnt_d[min((int)dep_d[t - m_ll], 15)], 1);
        }
    }
    __syncthreads();
    if (tid == 0) kraft_repair(cnt_ll, 15); else if (tid == 32) kraft_repair(cnt_d, 15);
    __syncthreads();
    // ---- lengths: the most frequent symbols get the shortest codes (sorted index j from the top)
    for (int t = tid; t < m_ll + m_d; t += kBuildThreads) {
        const bool isd = t >= m_ll;
        const int m = isd ? m_d : m_ll, j = isd ? t - m_ll : t;
        const int* cnt = isd ? cnt_d : cnt_ll;
        const int top = m - 1 - j;
        int l = 1, acc = cnt[1];
        while (acc <= top && l < 15) { l++; acc += cnt[l]; }
        if (isd) lens[kNumLL + order_d[j]] = (uint8_t)l; else lens[order_ll[j]] = (uint8_t)l;
    }
    __syncthreads();
    // ---- canonical codes (parallel): code = first code of that length + rank among equal lengths
    for (int t = tid; t < kHistSize; t += kBuildThreads) {
        const bool isd = t >= kNumLL;
        const int lo = isd ? kNumLL : 0, hi = isd ? kHistSize : kNumLL;
        const int* cnt = isd ? cnt_d : cnt_ll;
        const int l = lens[t];
        if (l) {
            int rank = 0;
            for (int j = lo; j < t; j++) rank += lens[j] == l;
            uint32_t c = 0;
            for (int k = 1; k <= l; k++) c = (c + (k > 1 ? cnt[k - 1] : 0)) << 1;
            codes[t] = (uint16_t)bitrev(c + rank, l);
        } else codes[t] = 0;
        int extra;
        if (!isd) extra = t >= 257 ? len_extra(t - 257) : 0; else extra = dist_extra(t - kNumLL);
        cost[t] = (uint32_t)l + (uint32_t)extra;
    }
    __syncthreads();
    // ---- header (serial) — and, in other warps, the exact bit size of every sub-chunk
    if (tid == 0) {
        int nll = kNumLL; while (nll > 257 && lens[nll - 1] == 0) nll--;
        int nd = kNumD; while (nd > 1 && lens[kNumLL + nd - 1] == 0) nd--;
        uint16_t* rl = reinterpret_cast<uint16_t*>(w_ll);      // w_ll is dead now (>= 572 u32 = 1144 u16 entries)
        int nrl = rle_lengths(lens, nll, rl);
        nrl += rle_lengths(lens + kNumLL, nd, rl + nrl);
        uint32_t clf[19]; uint8_t cll[19]; uint16_t clc[19]; uint16_t clo[19];
        for (int i = 0; i < 19; i++) { clf[i] = 0; cll[i] = 0; }
        for (int i = 0; i < nrl; i++) clf[rl[i] & 0xFF]++;
        int used = 0; for (int i = 0; i < 19; i++) used += clf[i] != 0;
        if (used == 0) { clf[0] = 1; clf[1] = 1; }
        else if (used == 1) { if (clf[0]) clf[1] = 1; else clf[0] = 1; }
        int m = 0;
        for (int i = 0; i < 19; i++) if (clf[i]) {       // insertion sort by (freq, sym)
            int p = m++;
            while (p > 0 && clf[clo[p - 1]] > clf[i]) { clo[p] = clo[p - 1]; p--; }
            clo[p] = (uint16_t)i;
        }
        uint32_t wcl[38]; uint16_t pcl[38];
        tree_serial(clf, clo, m, 7, cll, wcl, pcl);
        canonical_serial(cll, 19, 7, clc);
        int ncl = 19; while (ncl > 4 && cll[kClOrder[ncl - 1]] == 0) ncl--;
        BitW bw{B.blk_hdr + (size_t)b * kHdrBytes, 0, 0, 0};
        bw.put(blk.last ? 1u : 0u, 1); bw.put(2u, 2);
        bw.put(nll - 257, 5); bw.put(nd - 1, 5); bw.put(ncl - 4, 4);
        for (int i = 0; i < ncl; i++) bw.put(cll[kClOrder[i]], 3);
        for (int i = 0; i < nrl; i++) {
            const int s = rl[i] & 0xFF, x = rl[i] >> 8;
            bw.put(clc[s], cll[s]);
            if (s == 16) bw.put(x, 2); else if (s == 17) bw.put(x, 3); else if (s == 18) bw.put(x, 7);
        }
        bw.flush();
        hdr_bits_s = bw.total;
    } else if (tid >= 32) {
        const int warp = (tid - 32) >> 5, lane = tid & 31, nw = (kBuildThreads - 32) / 32;
        for (int j = warp; j < blk.nsub; j += nw) {
            const uint32_t* h = B.sub_hist + (size_t)(blk.sub0 + j) * kHistSize;
            unsigned long long s = 0;
            for (int i = lane; i < kHistSize; i += 32) s += (unsigned long long)h[i] * cost[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) sub_bits[j] = s;
        }
    }
    __syncthreads();
    // ---- publish codes, offsets, sizes
    for (int t = tid; t < kCodeStride; t += kBuildThreads) {
        B.blk_code[(size_t)b * kCodeStride + t] = codes[t];
        B.blk_clen[(size_t)b * kCodeStride + t] = lens[t];
    }
    if (tid == 0) {
        unsigned long long off = hdr_bits_s;
        for (int j = 0; j < blk.nsub; j++) { B.sub_bitoff[blk.sub0 + j] = off; off += sub_bits[j]; }
        const unsigned long long eob = off;
        unsigned long long bits = off + lens[256] + (blk.last ? 0 : 3);
        const unsigned long long hbytes = (bits + 7) / 8 + (blk.last ? 0 : 4);
        const unsigned long long rawlen = (unsigned long long)blk.len;
        unsigned long long sbytes = rawlen + 5ull * ((rawlen + 65534ull) / 65535ull);
        if (rawlen == 0) sbytes = 5;
        const bool stored = (B.level == 0) || (sbytes <= hbytes);
        const unsigned long long body = stored ? sbytes : hbytes;
        B.blk_hdr_bits[b] = hdr_bits_s;
        B.blk_eob_bit[b] = eob;
        B.blk_body_bits[b] = bits;
        B.blk_stored[b] = stored ? 1u : 0u;
        B.blk_len[b] = (uint32_t)((blk.first ? 2 : 0) + body + (blk.last ? 4 : 0));
    }
}

int launch_huff_build(const BatchD& b, cudaStream_t st) {
    if (b.nblocks == 0) return 0;
    k_huff_build<<<b.nblocks, kBuildThreads, 0, st>>>(b);
    return 1;
}

// ------------------------------------------------------------------------------------------ layout
// One CTA: exclusive scan over blocks (in page order) of chunk sizes -> positions; page extents.
// Framed layout of a page:  sig(8) IHDR(25) { len(4) "IDAT"(4) payload crc(4) }* IEND(12); page starts 16-byte aligned.
__global__ void __launch_bounds__(1024) k_layout(BatchD B) {
    __shared__ unsigned long long wsum[32];
    __shared__ unsigned long long carry_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // pass 1: per-page PNG length (pages are few thousand at most; blocks of a page are contiguous)
    for (int p = tid; p < B.npages; p += blockDim.x) {
        const PageD& P = B.pages[p];
        unsigned long long n = B.framed ? 8 + 25 + 12 : 0;
        for (int k = 0; k < P.nblk; k++) n += (unsigned long long)B.blk_len[P.blk0 + k] + (B.framed ? 12 : 0);
        B.png_len[p] = n;
        B.b64_len[p] = B.want_b64 ? 4ull * ((n + 2) / 3) : 0ull;
    }
    if (tid == 0) carry_s = 0;
    __syncthreads();
    // pass 2: exclusive scan of 16-byte-aligned page extents (png and b64 separately, two rounds)
    for (int round = 0; round < 2; round++) {
        uint64_t* len = round == 0 ? B.png_len : B.b64_len;
        uint64_t* off = round == 0 ? B.png_off : B.b64_off;
        if (tid == 0) carry_s = 0;
        __syncthreads();
        for (int base = 0; base < B.npages; base += blockDim.x) {
            const int p = base + tid;
            unsigned long long v = p < B.npages ? ((len[p] + 15ull) & ~15ull) : 0ull;
            unsigned long long x = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
            if (lane == 31) wsum[warp] = x;
            __syncthreads();
            if (warp == 0) {
                unsigned long long s = wsum[lane];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const unsigned long long y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
                wsum[lane] = s;
            }
            __syncthreads();
            const unsigned long long before = carry_s + (warp ? wsum[warp - 1] : 0ull) + (x - v);
            if (p < B.npages) off[p] = before;
            __syncthreads();
            if (tid == blockDim.x - 1) carry_s = before + v;
            __syncthreads();
        }
        if (tid == 0) {
            B.totals[round] = carry_s;
            if (carry_s > (round == 0 ? B.png_cap : B.b64_cap)) atomicOr(B.err, 1u << round);
        }
        __syncthreads();
    }
    // pass 3: payload positions
    for (int p = tid; p < B.npages; p += blockDim.x) {
        const PageD& P = B.pages[p];
        unsigned long long o = B.png_off[p] + (B.framed ? 8 + 25 : 0);
        for (int k = 0; k < P.nblk; k++) {
            if (B.framed) o += 8;
            B.blk_dst[P.blk0 + k] = o;
            o += B.blk_len[P.blk0 + k];
            if (B.framed) o += 4;
        }
    }
}

int launch_layout(const BatchD& b, cudaStream_t st) {
    if (b.npages == 0) return 0;
    k_layout<<<1, 1024, 0, st>>>(b);
    return 1;
}

// ------------------------------------------------------------------------------------------ payload init
__global__ void __launch_bounds__(256) k_payload_init(BatchD B) {
    if (B.err[0]) return;
    const int b = blockIdx.x, tid = threadIdx.x;
    const BlockD blk = B.blocks[b];
    const PageD& P = B.pages[blk.page];
    uint8_t* pay = B.png + B.blk_dst[b];
    const uint32_t plen = B.blk_len[b];
    uint8_t* body = pay + (blk.first ? 2 : 0);
    if (tid == 0 && blk.first) { pay[0] = 0x78; pay[1] = 0x9C; }
    if (tid == 0 && blk.last) {
        const uint32_t ad = B.page_adler[blk.page];
        uint8_t* t = pay + plen - 4;
        t[0] = (uint8_t)(ad >> 24); t[1] = (uint8_t)(ad >> 16); t[2] = (uint8_t)(ad >> 8); t[3] = (uint8_t)ad;
    }
    if (B.blk_stored[b]) {
        // stored blocks of <= 65535 bytes: [fin][len lo][len hi][~len lo][~len hi] data
        const uint8_t* raw = P.filt + blk.start;
        const long long rawlen = blk.len;
        const long long nsb = rawlen == 0 ? 1 : (rawlen + 65534) / 65535;
        for (long long k = tid; k < nsb; k += blockDim.x) {
            const long long off = k * 65535;
            const long long n = min(65535ll, rawlen - off);
            uint8_t* h = body + off + 5 * k;
            h[0] = (uint8_t)((blk.last && k == nsb - 1) ? 1 : 0);
            h[1] = (uint8_t)(n & 255); h[2] = (uint8_t)(n >> 8);
            h[3] = (uint8_t)(~n & 255); h[4] = (uint8_t)((~n >> 8) & 255);
        }
        for (long long i = tid; i < rawlen; i += blockDim.x) body[i + 5 * (i / 65535 + 1)] = raw[i];
        return;
    }
    // Huffman block: zero the body, then header bits, EOB, sync marker
    const unsigned long long bits = B.blk_body_bits[b];
    const long long body_bytes = (long long)((bits + 7) / 8) + (blk.last ? 0 : 4);
    {   // zero [body, body + body_bytes): bytes up to 4-byte alignment, then words
        const int head = (int)((4 - ((uintptr_t)body & 3)) & 3);
        const long long nh = min((long long)head, body_bytes);
        if (tid < nh) body[tid] = 0;
        uint32_t* w = reinterpret_cast<uint32_t*>(body + nh);
        const long long nw = (body_bytes - nh) >> 2;
        for (long long i = tid; i < nw; i += blockDim.x) w[i] = 0u;
        const long long tail0 = nh + 4 * nw;
        if (tail0 + tid < body_bytes && tid < 4) body[tail0 + tid] = 0;
    }
    __syncthreads();
    const uint32_t hb = B.blk_hdr_bits[b];
    const uint8_t* hsrc = B.blk_hdr + (size_t)b * kHdrBytes;
    for (int i = tid; i < (int)((hb + 7) / 8); i += blockDim.x) body[i] = hsrc[i];
    __syncthreads();
    if (tid == 0) {
        const unsigned long long eob = B.blk_eob_bit[b];
        const uint32_t code = B.blk_code[(size_t)b * kCodeStride + 256];
        const int l = B.blk_clen[(size_t)b * kCodeStride + 256];
        unsigned long long v = (unsigned long long)code << (eob & 7);
        long long byte = (long long)(eob >> 3);
        for (int n = l + (int)(eob & 7); n > 0; n -= 8) { body[byte++] |= (uint8_t)v; v >>= 8; }
        if (!blk.last) {
            uint8_t* m = body + (bits + 7) / 8;
            m[0] = 0; m[1] = 0; m[2] = 0xFF; m[3] = 0xFF;
        }
    }
}

int launch_payload_init(const BatchD& b, cudaStream_t st) {
    if (b.nblocks == 0) return 0;
    k_payload_init<<<b.nblocks, 256, 0, st>>>(b);
    return 1;
}

// ------------------------------------------------------------------------------------------ emit
__global__ void __launch_bounds__(kEmitThreads) k_huff_emit(BatchD B) {
    if (B.err[0]) return;
    __shared__ uint16_t s_code[kCodeStride];
    __shared__ uint8_t s_len[kCodeStride];
    __shared__ uint32_t buf[kEmitThreads * 48 / 32 + 4];
    __shared__ uint32_t wtot[kEmitThreads / 32];
    const int sub = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = (int)B.sub2blk[sub];
    if (B.blk_stored[b]) return;
    const BlockD blk = B.blocks[b];
    const PageD& P = B.pages[blk.page];
    for (int i = tid; i < kCodeStride; i += kEmitThreads) {
        s_code[i] = B.blk_code[(size_t)b * kCodeStride + i];
        s_len[i] = B.blk_clen[(size_t)b * kCodeStride + i];
    }
    const uint32_t ntok = B.sub_ntok[sub];
    const long long s = blk.start + (long long)(sub - blk.sub0) * kSubBytes;
    const uint32_t* __restrict__ tok = B.tokens + ((P.filt - B.filt_base) + s);
    // absolute bit position inside the png buffer (B.png is 256-byte aligned)
    unsigned long long pos = 8ull * (B.blk_dst[b] + (blk.first ? 2 : 0)) + B.sub_bitoff[sub];
    uint32_t* __restrict__ gw = reinterpret_cast<uint32_t*>(B.png);
    const unsigned long long first_word = pos >> 5;
    uint32_t carry = 0;                                   // bits of the current partial word (thread 0's copy is used)
    constexpr int kBufWords = kEmitThreads * 48 / 32 + 4;
    __syncthreads();
    for (uint32_t t0 = 0; t0 < ntok; t0 += kEmitThreads) {
        // ---- encode
        unsigned long long bits = 0; int nb = 0;
        if (t0 + tid < ntok) {
            const uint32_t k = __ldg(tok + t0 + tid);
            if (k & 0x80000000u) {
                const int L = (int)(k & 0xFFu) + 3, d = (int)((k >> 8) & 0x7FFFu) + 1;
                const int ls = len_sym(L), ds = dist_sym(d);
                const int le = len_extra(ls), de = dist_extra(ds);
                const int l1 = s_len[257 + ls], l2 = s_len[kNumLL + ds];
                bits = s_code[257 + ls];
                nb = l1;
                bits |= (unsigned long long)((uint32_t)(L - 3) & ((1u << le) - 1u)) << nb; nb += le;
                bits |= (unsigned long long)s_code[kNumLL + ds] << nb; nb += l2;
                bits |= (unsigned long long)((uint32_t)(d - 1) & ((1u << de) - 1u)) << nb; nb += de;
            } else {
                bits = s_code[k]; nb = s_len[k];
            }
        }
        // ---- block-wide exclusive scan of nb
        int x = nb;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) wtot[warp] = (uint32_t)x;
        for (int i = tid; i < kBufWords; i += kEmitThreads) buf[i] = 0u;
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int wi = 0; wi < kEmitThreads / 32; wi++) { const int v = (int)wtot[wi]; if (wi < warp) before += v; total += v; }
        const int sh0 = (int)(pos & 31);                  // tile starts at bit sh0 of buf[0]
        if (tid == 0) buf[0] = carry;
        __syncthreads();
        if (nb) {
            const int st = sh0 + before + (x - nb);
            const int wi = st >> 5, sh = st & 31;
            const unsigned long long lo = bits << sh;     // nb <= 48, sh <= 31: up to 79 bits -> three words
            atomicOr(&buf[wi], (uint32_t)lo);
            const uint32_t mid = (uint32_t)(lo >> 32);
            if (mid) atomicOr(&buf[wi + 1], mid);
            if (sh && nb + sh > 64) { const uint32_t hi = (uint32_t)(bits >> (64 - sh)); if (hi) atomicOr(&buf[wi + 2], hi); }
        }
        __syncthreads();
        // ---- flush complete words
        const int endbit = sh0 + total;
        const int nfull = endbit >> 5;
        const unsigned long long w0 = pos >> 5;
        for (int i = tid; i < nfull; i += kEmitThreads) {
            const uint32_t v = buf[i];
            if (w0 + i == first_word) atomicOr(&gw[w0 + i], v); else gw[w0 + i] = v;
        }
        if (tid == 0) carry = (endbit & 31) ? buf[nfull] : 0u;
        pos += (unsigned long long)total;
        __syncthreads();
    }
    if (tid == 0 && (pos & 31) && carry) atomicOr(&gw[pos >> 5], carry);
}

int launch_huff_emit(const BatchD& b, cudaStream_t st) {
    if (b.nsub == 0) return 0;
    k_huff_emit<<<b.nsub, kEmitThreads, 0, st>>>(b);
    return 1;
}

}  // namespace vcp
