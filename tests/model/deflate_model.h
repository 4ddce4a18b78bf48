/* deflate_model.h — CPU model of the GPU deflate kernels (test infrastructure, see deflate_model.c). */
#ifndef DEFLATE_MODEL_H
#define DEFLATE_MODEL_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct {
    int32_t bpp;          /* bytes per pixel of the filtered rows (second run distance) */
    int32_t hash_bits;    /* per-warp hash table: 2^hash_bits buckets */
    int32_t ways;         /* entries per bucket */
    int32_t lane_cap;     /* exact compare length of a hash candidate inside a lane */
    int32_t too_far;      /* length-3 matches farther than this are dropped */
    int32_t lazy;         /* 0 = greedy; else lane i yields if len[i] < lazy and len[i+1] > len[i] */
    int32_t cont_min;     /* after a 258 match continue at the same distance if >= cont_min bytes match (0 = off) */
    int32_t prime_bytes;  /* history inserted into the table before the sub-chunk (<= 32768) */
    int32_t capped_wins;  /* 1: a candidate that hit its compare cap outranks any exact one */
    int32_t inwin;        /* 1: nearest earlier same-hash lane of the window is a candidate */
    int32_t cont_maxd;    /* continuation only for distances <= cont_maxd */
    int32_t sub_bytes;    /* sub-chunk (one warp) */
    int32_t hash2_bytes;  /* 0 = off; else a second 1-way table keyed by a hash of this many bytes */
    int32_t hash2_bits;
    int32_t noisy_thresh; /* 0 = off; window is 'noisy' when the literal EMA (8x literals/window) >= this */
    int32_t noisy_minlen; /* in noisy windows matches shorter than this (and farther than noisy_neard) are dropped */
    int32_t noisy_neard;
    int32_t cost_maxlen;  /* 0 = off; matches up to this length are priced against literals with the running histogram */
    int32_t cost_margin;  /* quarter bits a match must save */
    int32_t cost_warm;    /* tokens needed in the sub-chunk before pricing starts */
    int32_t hash2_ways;
    int32_t ins_limit;    /* 1: a window inserts only the positions its parse consumed (q < next) */
    int32_t noisy_ways1;  /* 1: in noisy windows only the newest way of each table is a candidate */
    int32_t lane_cap_win; /* 1: lanes 0-15 compare candidates up to 32 bytes, lanes 16-31 up to 16 */
    int32_t rowlen;       /* > 0: extra candidate at distance rowlen (one filtered row up) */
    int32_t row_gate;     /* 0: always probe; 1: only where S[q] != S[q-1]; 2: not in noisy windows */
    int32_t exact_sel;    /* 1: a selected token whose lane has capped candidates compares all of them to the limit */
    int32_t group_subs;   /* consecutive sub-chunks of a page that share one pair of tables (first primed, rest continue) */
    int32_t prime_win;    /* positions per priming step (32 or 128): within a step only the highest position of a bucket is inserted */
    int32_t cost_epoch;   /* 0: costs from the counts as of the window start; N: from a table refreshed at the first window after every N tokens */
    int64_t block_bytes;  /* deflate block (multiple of sub_bytes) */
} dm_params;
typedef struct { int64_t tokens, blocks, stored_blocks; } dm_stats;
int dm_len_sym(int len);
int dm_dist_sym(int dist);
int64_t dm_lz_subchunk(const uint8_t* S, int64_t F, int64_t s, int64_t e, const dm_params* P, uint32_t* tok, uint32_t* hist);
int64_t dm_lz_subchunk_ex(const uint8_t* S, int64_t F, int64_t s, int64_t e, const dm_params* P, uint32_t* tok, uint32_t* hist,
                          uint16_t* T, uint16_t* T2, int cont);
void dm_huff_lengths(const uint32_t* freq, int n, int maxbits, uint8_t* lens);
void dm_canonical(const uint8_t* lens, int n, int maxbits, uint16_t* codes);
int64_t dm_huff_block(const uint32_t* tok, int64_t ntok, const uint32_t* hist, const uint8_t* raw, int64_t rawlen,
                      int first, int last, uint32_t adler, uint8_t* out, int* used_stored);
int64_t dm_deflate_page(const uint8_t* S, int64_t F, const dm_params* P, uint8_t* out, int64_t* block_sizes, dm_stats* st);
#ifdef __cplusplus
}
#endif
#endif
