/* deflate_model.c — sequential CPU model of the GPU deflate (TEST INFRASTRUCTURE).
 *
 * This is NOT the reference's algorithm (that is zlib, called by Pillow); it is a
 * lane-by-lane restatement of what vision_compression_project_b200/csrc/deflate_*.cu
 * computes, so that (a) compression-ratio decisions can be explored on the CPU and
 * (b) the GPU's token stream / byte stream can be compared bit-for-bit in tests.
 * Validity of the produced stream is checked independently with zlib's inflate.
 *
 * Structure mirrored from the kernels:
 *   page stream (filtered PNG rows) -> deflate blocks of `block_bytes`
 *                                   -> sub-chunks of `sub_bytes` (one warp each)
 *   sub-chunk: 32-position windows; every lane proposes a match (hash candidate(s),
 *   distance-1 run, distance-bpp run), the greedy parse is resolved from lane 0, the
 *   last token of a window may be extended cooperatively up to 258.
 *   block: one dynamic-Huffman block over all its sub-chunks' tokens, stored fallback,
 *   byte-aligned with an empty stored block so blocks concatenate at byte granularity.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#include "deflate_model.h"

#define MAXD 32768
#define MAXLEN 258
#define WIN 32

static const uint16_t LEN_BASE[29] = {3,4,5,6,7,8,9,10,11,13,15,17,19,23,27,31,35,43,51,59,67,83,99,115,131,163,195,227,258};
static const uint8_t LEN_EXTRA[29] = {0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0};
static const uint16_t DIST_BASE[30] = {1,2,3,4,5,7,9,13,17,25,33,49,65,97,129,193,257,385,513,769,1025,1537,2049,3073,4097,6145,8193,12289,16385,24577};
static const uint8_t DIST_EXTRA[30] = {0,0,0,0,1,1,2,2,3,3,4,4,5,5,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13};
static const uint8_t CL_ORDER[19] = {16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15};

int dm_len_sym(int len) {           /* len 3..258 -> 0..28 */
    int s = 28;
    if (len == 258) return 28;
    for (s = 0; s < 28; s++) if (len < LEN_BASE[s + 1]) break;
    return s;
}
int dm_dist_sym(int dist) {         /* dist 1..32768 -> 0..29 */
    int s;
    for (s = 0; s < 29; s++) if (dist < DIST_BASE[s + 1]) break;
    return s;
}

static inline uint32_t hash3(const uint8_t* p, int hb) {
    uint32_t v = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
    return (v * 0x9E3779B1u) >> (32 - hb);
}

static inline uint32_t hashN(const uint8_t* p, int nb, int hb) {
    /* 32-bit mix of up to 8 bytes: low word and high word multiplied by different odd constants */
    uint32_t lo = 0, hi = 0;
    for (int k = 0; k < nb && k < 4; k++) lo |= (uint32_t)p[k] << (8 * k);
    for (int k = 4; k < nb; k++) hi |= (uint32_t)p[k] << (8 * (k - 4));
    return (lo * 0x9E3779B1u + hi * 0x85EBCA77u) >> (32 - hb);
}

static inline int match_len(const uint8_t* a, const uint8_t* b, int cap) {
    int n = 0;
    while (n < cap && a[n] == b[n]) n++;
    return n;
}

/* quarter-bit integer log2: 4*floor(log2 v) + the two mantissa bits below the leading one (v >= 1) */
static inline int ilog2x4(uint32_t v) {
    int n = 31 - __builtin_clz(v);
    uint32_t frac = n >= 2 ? (v >> (n - 2)) & 3u : (n == 1 ? (v & 1u) << 1 : 0);
    return 4 * n + (int)frac;
}

/* ------------------------------------------------------------------ LZ: one sub-chunk (one warp) */
/* T / T2: caller-owned tables (T: ways << hash_bits, T2: hash2_ways << hash2_bits u16 entries).
   cont = 0: tables are cleared and primed with the previous prime_bytes of the stream (independent sub-chunk);
   cont = 1: tables hold the state left by the sub-chunk [s - sub_bytes, s) (entries relative to s - sub_bytes - 32 KiB);
             they are rebased by sub_bytes (entries that fall out of the window become empty) and not primed. */
int64_t dm_lz_subchunk_ex(const uint8_t* S, int64_t F, int64_t s, int64_t e, const dm_params* P,
                          uint32_t* tok, uint32_t* hist /* 286 + 30 */, uint16_t* T, uint16_t* T2, int cont) {
    const int hb = P->hash_bits, ways = P->ways, bpp = P->bpp, lcap = P->lane_cap;
    const int64_t base = s - MAXD;                  /* table stores pos - base as u16 */
    const int tsize = (1 << hb) * ways;
    const int tsize2 = P->hash2_bytes ? ((P->hash2_ways > 0 ? P->hash2_ways : 1) << P->hash2_bits) : 0;
    if (!cont) {
        memset(T, 0, sizeof(uint16_t) * tsize);   /* 0 = empty (position base+0 is never a candidate) */
        if (tsize2) memset(T2, 0, sizeof(uint16_t) * tsize2);
    } else {
        for (int i = 0; i < tsize; i++) T[i] = T[i] >= P->sub_bytes ? (uint16_t)(T[i] - P->sub_bytes) : 0;
        for (int i = 0; i < tsize2; i++) T2[i] = T2[i] >= P->sub_bytes ? (uint16_t)(T2[i] - P->sub_bytes) : 0;
    }
    int64_t ntok = 0;
    memset(hist, 0, sizeof(uint32_t) * 316);
    const int nb2 = P->hash2_bytes, hb2 = P->hash2_bits;
    const int ways2 = P->hash2_ways > 0 ? P->hash2_ways : 1;

#define INSERT(q) do { if ((q) >= 0 && (q) + 2 < F) { uint32_t h_ = hash3(S + (q), hb) * ways; \
        for (int w_ = ways - 1; w_ > 0; w_--) { T[h_ + w_] = T[h_ + w_ - 1]; } \
        T[h_] = (uint16_t)((q) - base); } } while (0)
#define INSERT2(q) do { if (nb2 && (q) >= 0 && (q) + nb2 <= F) { uint32_t h_ = hashN(S + (q), nb2, hb2) * ways2; \
        for (int w_ = ways2 - 1; w_ > 0; w_--) { T2[h_ + w_] = T2[h_ + w_ - 1]; } \
        T2[h_] = (uint16_t)((q) - base); } } while (0)

    /* priming: positions [max(0, s - prime), s), window by window; within a window only the
       highest lane of each hash group writes, and for ways>1 it shifts the bucket ONCE. */
    int64_t hs = s - P->prime_bytes; if (hs < 0) hs = 0;
    if (cont) hs = s;                               /* continuing: nothing to prime */
    const int PW = P->prime_win > 0 ? P->prime_win : WIN;
    for (int64_t w0 = hs; w0 < s; w0 += PW) {
        for (int i = 0; i < PW; i++) {
            int64_t q = w0 + i; if (q >= s) continue;
            /* per table: only the highest position of a same-bucket group writes (atomicMax on the device) */
            if (q + 2 < F) {
                uint32_t h = hash3(S + q, hb); int winner = 1;
                for (int j = i + 1; j < PW; j++) { int64_t q2 = w0 + j; if (q2 >= s || q2 + 2 >= F) continue; if (hash3(S + q2, hb) == h) { winner = 0; break; } }
                if (winner) INSERT(q);
            }
            if (nb2 && q + nb2 <= F) {
                uint32_t h = hashN(S + q, nb2, hb2); int winner = 1;
                for (int j = i + 1; j < PW; j++) { int64_t q2 = w0 + j; if (q2 >= s || q2 + nb2 > F) continue; if (hashN(S + q2, nb2, hb2) == h) { winner = 0; break; } }
                if (winner) INSERT2(q);
            }
        }
    }

    int64_t p = s;
    int score = 0;                     /* EMA of literal tokens per window (x8) */
    uint32_t hsnap[316]; uint32_t Nsnap = 0; int64_t epoch = 0;
    memset(hsnap, 0, sizeof hsnap);
    while (p < e) {
        const int noisy = P->noisy_thresh < 0 || (P->noisy_thresh > 0 && score >= P->noisy_thresh);   /* < 0: always */
        /* costs use the counts as of the window start, or (cost_epoch) as of the first window after every cost_epoch tokens */
        if (!P->cost_epoch || ntok / P->cost_epoch != epoch) { memcpy(hsnap, hist, sizeof hsnap); Nsnap = (uint32_t)ntok; if (P->cost_epoch) epoch = ntok / P->cost_epoch; }
        uint32_t Ntok = (uint32_t)ntok;
        int nlit = 0;
        int len[WIN], dist[WIN], capped[WIN];
        /* exact_sel: per lane, the capped candidates (distance, length so far) and the best fully compared one */
        int ccd[WIN][8], ccl[WIN][8], ncc[WIN], xl[WIN], xd[WIN];
        /* ---- every lane proposes */
        for (int i = 0; i < WIN; i++) {
            int64_t q = p + i; len[i] = 0; dist[i] = 0; capped[i] = 0; ncc[i] = 0; xl[i] = 0; xd[i] = 0;
            if (q >= e) continue;
            int limit = (int)((e - q) < MAXLEN ? (e - q) : MAXLEN);
            if (limit < 3) continue;
            int bl = 0, bd = 0, bc = 0, be = 0;
            /* distance-1 and distance-bpp runs: exact up to the end of a 64-position mask that
               starts at the window base (positions p .. p+63) */
            int runcap = 64 - i; if (runcap > limit) runcap = limit;
            int dd[2] = {1, bpp};
            for (int k = 0; k < (bpp > 1 ? 2 : 1); k++) {
                int d = dd[k];
                if (q - d < 0) continue;
                int l = match_len(S + q, S + q - d, runcap);
                int c = (l == runcap && runcap < limit);
                int eff = (c && P->capped_wins) ? 1000 : l;
                if (l >= 3 && c) { ccd[i][ncc[i]] = d; ccl[i][ncc[i]] = l; ncc[i]++; }
                if (l >= 3 && !c && (l > xl[i] || (l == xl[i] && d < xd[i]))) { xl[i] = l; xd[i] = d; }
                if (l >= 3 && eff > be) { be = eff; bl = l; bd = d; bc = c; }
            }
            /* hash candidates, exact up to lane_cap: the most recent earlier lane of this window with
               the same hash (if any), then the table ways — at most `ways` candidates in total */
            if (q + 2 < F && !(bc && P->capped_wins)) {   /* a capped run already outranks everything */
                uint32_t hv = hash3(S + q, hb);
                uint32_t h = hv * ways;
                /* lane_cap_win: a lane only needs an exact length up to the end of its window (later lanes are cut to 16 bytes);
                   whatever is longer is 'capped' and, if it becomes the window's last token, extended cooperatively */
                const int lcap_i = P->lane_cap_win ? ((i < 16 ? 32 : 16) < lcap ? (i < 16 ? 32 : 16) : lcap) : lcap;
                int hcap = lcap_i < limit ? lcap_i : limit;
                int64_t cands[12]; int nc = 0;
                if (P->inwin) {
                    for (int j = i - 1; j >= 0; j--) {
                        int64_t q2 = p + j;
                        if (q2 + 2 < F && hash3(S + q2, hb) == hv) { cands[nc++] = q2; break; }
                    }
                }
                /* in a noisy stretch (literal EMA high) only the newest entry of each table is verified when noisy_ways1 is set */
                const int nw1 = (noisy && P->noisy_ways1) ? 1 : ways, nw2 = (noisy && P->noisy_ways1) ? 1 : ways2;
                for (int w = 0; w < nw1 && nc < ways; w++) {
                    uint16_t cnd = T[h + w]; if (cnd == 0) continue;
                    cands[nc++] = base + cnd;
                }
                if (nb2 && q + nb2 <= F) { uint32_t h2 = hashN(S + q, nb2, hb2) * ways2; for (int w = 0; w < nw2; w++) { uint16_t cnd = T2[h2 + w]; if (cnd != 0) cands[nc++] = base + cnd; } }
                /* row probe: the byte one filtered row above (distance = 1 + width*bpp), where smooth shading repeats exactly */
                if (P->rowlen > 0 && q - P->rowlen >= 0 && P->rowlen <= MAXD &&
                    !(P->row_gate == 1 && q >= 1 && S[q] == S[q - 1]) && !(P->row_gate == 2 && noisy)) cands[nc++] = q - P->rowlen;
                for (int w = 0; w < nc; w++) {
                    int64_t cp = cands[w];
                    if (cp >= q || q - cp > MAXD) continue;
                    int l = match_len(S + q, S + cp, hcap);
                    int c = (l == hcap && hcap < limit);
                    int d = (int)(q - cp);
                    int eff = (c && P->capped_wins) ? 1000 : l;
                    if (l >= 3 && c) { ccd[i][ncc[i]] = d; ccl[i][ncc[i]] = l; ncc[i]++; }
                    if (l >= 3 && !c && (l > xl[i] || (l == xl[i] && d < xd[i]))) { xl[i] = l; xd[i] = d; }
                    if (l >= 3 && (eff > be || (eff == be && d < bd))) { be = eff; bl = l; bd = d; bc = c; }
                }
            }
            if (bl < 3 || (bl == 3 && bd > P->too_far)) continue;
            if (noisy && bl < P->noisy_minlen && bd > P->noisy_neard) continue;
            if (P->cost_maxlen && bl <= P->cost_maxlen && Ntok >= (uint32_t)P->cost_warm) {
                int lgN = ilog2x4(Nsnap + 1);
                int lit = 0;
                for (int k = 0; k < bl; k++) lit += lgN - ilog2x4(hsnap[S[q + k]] + 1);
                int ls = dm_len_sym(bl), ds = dm_dist_sym(bd);
                int mc = (lgN - ilog2x4(hsnap[257 + ls] + 1)) + 4 * LEN_EXTRA[ls] + (lgN - ilog2x4(hsnap[286 + ds] + 1)) + 4 * DIST_EXTRA[ds];
                if (mc + P->cost_margin >= lit) continue;
            }
            len[i] = bl; dist[i] = bd; capped[i] = bc;
        }
        /* ---- optional one-step lazy rule, evaluated on all lanes at once */
        if (P->lazy) {
            int nl[WIN];
            for (int i = 0; i < WIN; i++) nl[i] = len[i];
            for (int i = 0; i + 1 < WIN; i++)
                if (len[i] >= 3 && len[i] < P->lazy && len[i + 1] > len[i]) nl[i] = 0;
            for (int i = 0; i < WIN; i++) if (!nl[i]) { len[i] = 0; }
        }
        /* ---- greedy parse from lane 0 */
        int i = 0; int64_t next = p;
        while (i < WIN && p + i < e) {
            int64_t q = p + i;
            if (len[i] >= 3) {
                int L = len[i], d = dist[i];
                if (capped[i] && P->exact_sel) {
                    /* a selected token with capped candidates: compare every one of them (and the best exact one) to the limit */
                    int limit = (int)((e - q) < MAXLEN ? (e - q) : MAXLEN);
                    L = xl[i]; d = xd[i];
                    for (int k = 0; k < ncc[i]; k++) {
                        int dk = ccd[i][k];
                        int Lk = ccl[i][k] + match_len(S + q + ccl[i][k], S + q + ccl[i][k] - dk, limit - ccl[i][k]);
                        if (Lk > L || (Lk == L && dk < d)) { L = Lk; d = dk; }
                    }
                } else if (capped[i]) {          /* cooperative extension: exact up to the limit */
                    int limit = (int)((e - q) < MAXLEN ? (e - q) : MAXLEN);
                    L += match_len(S + q + L, S + q + L - d, limit - L);
                }
                tok[ntok++] = 0x80000000u | ((uint32_t)(d - 1) << 8) | (uint32_t)(L - 3);
                hist[257 + dm_len_sym(L)]++; hist[286 + dm_dist_sym(d)]++;
                i += L; next = q + L;
                /* continuation: after a maximal match keep going at the same distance */
                if (L == MAXLEN && P->cont_min > 0 && d <= P->cont_maxd) {
                    for (;;) {
                        int64_t q2 = next; if (q2 >= e) break;
                        int limit = (int)((e - q2) < MAXLEN ? (e - q2) : MAXLEN);
                        int l2 = match_len(S + q2, S + q2 - d, limit);
                        if (l2 < P->cont_min) break;
                        tok[ntok++] = 0x80000000u | ((uint32_t)(d - 1) << 8) | (uint32_t)(l2 - 3);
                        hist[257 + dm_len_sym(l2)]++; hist[286 + dm_dist_sym(d)]++;
                        next = q2 + l2; i = WIN;
                        if (l2 < MAXLEN) break;
                    }
                }
            } else {
                tok[ntok++] = S[q]; hist[S[q]]++; nlit++;
                i += 1; next = q + 1;
            }
        }
        /* ---- insert the window's positions (highest lane of a same-bucket group wins); with ins_limit only the
           positions the parse has consumed (q < next) are inserted, so a position is never inserted twice */
        const int64_t iend = (P->ins_limit && next < e) ? next : e;
        for (int k = 0; k < WIN; k++) {
            int64_t q = p + k; if (q >= iend) continue;
            if (q + 2 < F) {
                uint32_t h = hash3(S + q, hb); int winner = 1;
                for (int j = k + 1; j < WIN; j++) { int64_t q2 = p + j; if (q2 >= iend || q2 + 2 >= F) continue; if (hash3(S + q2, hb) == h) { winner = 0; break; } }
                if (winner) INSERT(q);
            }
            if (nb2 && q + nb2 <= F) {
                uint32_t h = hashN(S + q, nb2, hb2); int winner = 1;
                for (int j = k + 1; j < WIN; j++) { int64_t q2 = p + j; if (q2 >= iend || q2 + nb2 > F) continue; if (hashN(S + q2, nb2, hb2) == h) { winner = 0; break; } }
                if (winner) INSERT2(q);
            }
        }
        p = next;
        score = score - (score >> 3) + nlit;
    }
    return ntok;
}

/* independent sub-chunk (fresh tables) */
int64_t dm_lz_subchunk(const uint8_t* S, int64_t F, int64_t s, int64_t e, const dm_params* P, uint32_t* tok, uint32_t* hist) {
    uint16_t* T = (uint16_t*)malloc(sizeof(uint16_t) * ((size_t)P->ways << P->hash_bits));
    uint16_t* T2 = (uint16_t*)malloc(sizeof(uint16_t) * (((size_t)(P->hash2_ways > 0 ? P->hash2_ways : 1) << P->hash2_bits) + 1));
    int64_t n = dm_lz_subchunk_ex(S, F, s, e, P, tok, hist, T, T2, 0);
    free(T); free(T2);
    return n;
}

/* ------------------------------------------------------------------ Huffman */
/* Length-limited code lengths for n symbols. Same steps as the kernel:
   force >= 2 used symbols, sort by (freq, sym), two-queue merge, leaf depths -> count per
   length, clamp to maxbits with a Kraft repair, hand lengths out in sorted order. */
void dm_huff_lengths(const uint32_t* freq_in, int n, int maxbits, uint8_t* lens) {
    uint32_t freq[288]; int used = 0;
    for (int i = 0; i < n; i++) { freq[i] = freq_in[i]; if (freq[i]) used++; lens[i] = 0; }
    if (used == 0) { freq[0] = 1; freq[1] = 1; used = 2; }
    else if (used == 1) { if (freq[0]) freq[1] = 1; else freq[0] = 1; used = 2; }
    int order[288]; int m = 0;
    for (int i = 0; i < n; i++) if (freq[i]) order[m++] = i;
    /* insertion sort by (freq asc, sym asc) — the kernel uses a rank sort with the same key */
    for (int a = 1; a < m; a++) { int v = order[a]; int b = a - 1;
        while (b >= 0 && (freq[order[b]] > freq[v] || (freq[order[b]] == freq[v] && order[b] > v))) { order[b + 1] = order[b]; b--; }
        order[b + 1] = v; }
    /* two-queue merge; nodes 0..m-1 leaves (sorted), m..2m-2 internal */
    uint64_t w[576]; int parent[576];
    for (int i = 0; i < m; i++) w[i] = freq[order[i]];
    int li = 0, ii = m, nn = m;
    while (nn < 2 * m - 1) {
        int pick[2];
        for (int k = 0; k < 2; k++) {
            if (li < m && (ii >= nn || w[li] <= w[ii])) pick[k] = li++;
            else pick[k] = ii++;
        }
        w[nn] = w[pick[0]] + w[pick[1]]; parent[pick[0]] = nn; parent[pick[1]] = nn; nn++;
    }
    int depth[576]; depth[2 * m - 2] = 0;
    for (int i = 2 * m - 3; i >= 0; i--) depth[i] = depth[parent[i]] + 1;
    int cnt[64]; memset(cnt, 0, sizeof cnt);
    for (int i = 0; i < m; i++) { int d = depth[i]; if (d > maxbits) d = maxbits; cnt[d]++; }
    /* Kraft repair in units of 2^-maxbits */
    int64_t K = 0; for (int l = 1; l <= maxbits; l++) K += (int64_t)cnt[l] << (maxbits - l);
    int64_t excess = K - ((int64_t)1 << maxbits);
    while (excess > 0) {
        int l = maxbits - 1; while (cnt[l] == 0) l--;
        cnt[l]--; cnt[l + 1]++; excess -= (int64_t)1 << (maxbits - l - 1);
    }
    while (excess < 0) {          /* slack: promote leaves where it fits exactly */
        int done = 0;
        for (int l = maxbits; l >= 2; l--) {
            int64_t cost = (int64_t)1 << (maxbits - l);
            if (cnt[l] > 0 && cost <= -excess) { cnt[l]--; cnt[l - 1]++; excess += cost; done = 1; break; }
        }
        if (!done) break;
    }
    /* most frequent symbols get the shortest lengths */
    int idx = m - 1;
    for (int l = 1; l <= maxbits; l++) for (int c = 0; c < cnt[l]; c++) lens[order[idx--]] = (uint8_t)l;
}

void dm_canonical(const uint8_t* lens, int n, int maxbits, uint16_t* codes /* bit-reversed */) {
    int cnt[17] = {0}; uint32_t next[17];
    for (int i = 0; i < n; i++) cnt[lens[i]]++;
    cnt[0] = 0; uint32_t c = 0;
    for (int l = 1; l <= maxbits; l++) { c = (c + cnt[l - 1]) << 1; next[l] = c; }
    for (int i = 0; i < n; i++) {
        int l = lens[i]; if (!l) { codes[i] = 0; continue; }
        uint32_t v = next[l]++, r = 0;
        for (int b = 0; b < l; b++) r |= ((v >> b) & 1u) << (l - 1 - b);
        codes[i] = (uint16_t)r;
    }
}

typedef struct { uint8_t* out; int64_t nbits; } bitw;
static inline void putbits(bitw* B, uint32_t v, int n) {
    for (int i = 0; i < n; i++) {
        int64_t b = B->nbits++;
        if ((b & 7) == 0) B->out[b >> 3] = 0;
        B->out[b >> 3] |= (uint8_t)(((v >> i) & 1u) << (b & 7));
    }
}

/* RLE of a code-length sequence into code-length-code symbols; returns count. sym|extra<<8 */
static int rle_lengths(const uint8_t* L, int n, uint16_t* out) {
    int k = 0, i = 0;
    while (i < n) {
        int v = L[i], r = 1; while (i + r < n && L[i + r] == v) r++;
        i += r;
        if (v == 0) {
            while (r >= 11) { int t = r > 138 ? 138 : r; out[k++] = (uint16_t)(18 | ((t - 11) << 8)); r -= t; }
            if (r >= 3) { out[k++] = (uint16_t)(17 | ((r - 3) << 8)); r = 0; }
            while (r-- > 0) out[k++] = 0;
        } else {
            out[k++] = (uint16_t)v; r--;
            while (r >= 3) { int t = r > 6 ? 6 : r; out[k++] = (uint16_t)(16 | ((t - 3) << 8)); r -= t; }
            while (r-- > 0) out[k++] = (uint16_t)v;
        }
    }
    return k;
}

/* One deflate block. tok = concatenated tokens of the block's sub-chunks, hist = summed histogram
   (without EOB). raw/rawlen = the block's bytes (stored fallback). Output is byte aligned.
   Returns bytes written. */
int64_t dm_huff_block(const uint32_t* tok, int64_t ntok, const uint32_t* hist_in, const uint8_t* raw, int64_t rawlen,
                      int first, int last, uint32_t adler, uint8_t* out, int* used_stored) {
    uint32_t hist[316]; memcpy(hist, hist_in, sizeof hist); hist[256] += 1;
    uint8_t ll_len[286], d_len[30], cl_len[19];
    uint16_t ll_code[286], d_code[30], cl_code[19];
    dm_huff_lengths(hist, 286, 15, ll_len);
    dm_huff_lengths(hist + 286, 30, 15, d_len);
    dm_canonical(ll_len, 286, 15, ll_code);
    dm_canonical(d_len, 30, 15, d_code);
    int nll = 286; while (nll > 257 && ll_len[nll - 1] == 0) nll--;
    int nd = 30; while (nd > 1 && d_len[nd - 1] == 0) nd--;
    uint16_t rl[320]; int nrl = rle_lengths(ll_len, nll, rl); nrl += rle_lengths(d_len, nd, rl + nrl);
    uint32_t clf[19] = {0}; for (int i = 0; i < nrl; i++) clf[rl[i] & 0xFF]++;
    dm_huff_lengths(clf, 19, 7, cl_len);
    dm_canonical(cl_len, 19, 7, cl_code);
    int ncl = 19; while (ncl > 4 && cl_len[CL_ORDER[ncl - 1]] == 0) ncl--;
    int64_t bits = 3 + 14 + 3 * ncl;
    for (int i = 0; i < nrl; i++) { int s = rl[i] & 0xFF; bits += cl_len[s] + (s == 16 ? 2 : s == 17 ? 3 : s == 18 ? 7 : 0); }
    for (int s = 0; s < 286; s++) bits += (int64_t)hist[s] * (ll_len[s] + (s >= 257 ? LEN_EXTRA[s - 257] : 0));
    for (int s = 0; s < 30; s++) bits += (int64_t)hist[286 + s] * (d_len[s] + DIST_EXTRA[s]);
    if (!last) bits += 3;
    int64_t hbytes = (bits + 7) / 8 + (last ? 0 : 4);
    int64_t sbytes = rawlen + 5 * ((rawlen + 65534) / 65535); if (rawlen == 0) sbytes = 5;
    int64_t o = 0;
    if (first) { out[o++] = 0x78; out[o++] = 0x9C; }
    if (sbytes <= hbytes) {
        *used_stored = 1;
        int64_t off = 0;
        do {
            int64_t n = rawlen - off; if (n > 65535) n = 65535;
            int fin = last && (off + n == rawlen);
            out[o++] = (uint8_t)fin; out[o++] = (uint8_t)(n & 255); out[o++] = (uint8_t)(n >> 8);
            out[o++] = (uint8_t)(~n & 255); out[o++] = (uint8_t)((~n >> 8) & 255);
            memcpy(out + o, raw + off, n); o += n; off += n;
        } while (off < rawlen);
    } else {
        *used_stored = 0;
        bitw B = {out + o, 0};
        putbits(&B, last ? 1 : 0, 1); putbits(&B, 2, 2);
        putbits(&B, nll - 257, 5); putbits(&B, nd - 1, 5); putbits(&B, ncl - 4, 4);
        for (int i = 0; i < ncl; i++) putbits(&B, cl_len[CL_ORDER[i]], 3);
        for (int i = 0; i < nrl; i++) {
            int s = rl[i] & 0xFF, x = rl[i] >> 8;
            putbits(&B, cl_code[s], cl_len[s]);
            if (s == 16) putbits(&B, x, 2); else if (s == 17) putbits(&B, x, 3); else if (s == 18) putbits(&B, x, 7);
        }
        for (int64_t t = 0; t < ntok; t++) {
            uint32_t k = tok[t];
            if (k & 0x80000000u) {
                int L = (int)(k & 0xFF) + 3, d = (int)((k >> 8) & 0x7FFF) + 1;
                int ls = dm_len_sym(L), ds = dm_dist_sym(d);
                putbits(&B, ll_code[257 + ls], ll_len[257 + ls]); putbits(&B, L - LEN_BASE[ls], LEN_EXTRA[ls]);
                putbits(&B, d_code[ds], d_len[ds]); putbits(&B, d - DIST_BASE[ds], DIST_EXTRA[ds]);
            } else putbits(&B, ll_code[k], ll_len[k]);
        }
        putbits(&B, ll_code[256], ll_len[256]);
        if (!last) putbits(&B, 0, 3);
        o += (B.nbits + 7) / 8;
        if (!last) { out[o++] = 0; out[o++] = 0; out[o++] = 0xFF; out[o++] = 0xFF; }
    }
    if (last) { out[o++] = adler >> 24; out[o++] = adler >> 16; out[o++] = adler >> 8; out[o++] = adler; }
    return o;
}

/* ------------------------------------------------------------------ whole page */
static uint32_t adler32_(const uint8_t* d, int64_t n) {
    uint32_t a = 1, b = 0;
    for (int64_t i = 0; i < n; i++) { a += d[i]; if (a >= 65521) a -= 65521; b += a; if (b >= 65521) b -= 65521; }
    return (b << 16) | a;
}

int64_t dm_deflate_page(const uint8_t* S, int64_t F, const dm_params* P, uint8_t* out, int64_t* block_sizes, dm_stats* st) {
    int64_t nblocks = (F + P->block_bytes - 1) / P->block_bytes; if (nblocks == 0) nblocks = 1;
    uint32_t adler = adler32_(S, F);
    uint32_t* tok = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(P->block_bytes + 64));
    uint16_t* T = (uint16_t*)malloc(sizeof(uint16_t) * ((size_t)P->ways << P->hash_bits));
    uint16_t* T2 = (uint16_t*)malloc(sizeof(uint16_t) * (((size_t)(P->hash2_ways > 0 ? P->hash2_ways : 1) << P->hash2_bits) + 1));
    const int G = P->group_subs > 1 ? P->group_subs : 1;      /* consecutive sub-chunks of a page handled by one warp */
    int64_t subidx = 0;
    int64_t o = 0;
    if (st) memset(st, 0, sizeof *st);
    for (int64_t b = 0; b < nblocks; b++) {
        int64_t bs = b * P->block_bytes, be = bs + P->block_bytes; if (be > F) be = F;
        uint32_t hist[316] = {0}, h1[316]; int64_t ntok = 0;
        for (int64_t s = bs; s < be; s += P->sub_bytes) {
            int64_t e = s + P->sub_bytes; if (e > be) e = be;
            ntok += dm_lz_subchunk_ex(S, F, s, e, P, tok + ntok, h1, T, T2, (int)(subidx % G != 0));
            subidx++;
            for (int i = 0; i < 316; i++) hist[i] += h1[i];
        }
        int stored = 0;
        int64_t n = dm_huff_block(tok, ntok, hist, S + bs, be - bs, b == 0, b == nblocks - 1, adler, out + o, &stored);
        if (block_sizes) block_sizes[b] = n;
        if (st) { st->tokens += ntok; st->stored_blocks += stored; st->blocks++; }
        o += n;
    }
    free(tok); free(T); free(T2);
    return o;
}
