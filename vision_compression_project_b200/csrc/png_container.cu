// png_container.cu — PNG chunk framing, CRC-32, Adler-32 and base64.
//
// Replaces the container writes of PIL/PngImagePlugin.py:1325-1525 (_save: signature, IHDR, IDAT via
// _idat/putchunk :1124-1143, IEND), zlib's crc32.c / adler32.c, and CPython's binascii.b2a_base64 behind
// base64.b64encode — i.e. everything after the deflate bits when the reference runs
// `page_image.save(path)` (backend/app/pipeline/pdf_extract.py:130) and the image->blob step of
// generate_content (pdf_extract.py:55).  Restated in oracle/restate.py (png_wrap, crc32, adler32, b64encode).
//
//   k_png_finish   CTA per IDAT: CRC-32 of "IDAT"+payload — every thread runs a table CRC over its own
//                  contiguous segment, segments are combined with the GF(2) identity
//                  crc(A||B) = crc(A) * x^(8|B|) mod P  xor  crc(B)  (no serial pass over the chunk);
//                  writes length/type/CRC; the page's first/last block also write sig+IHDR / IEND.
//   k_base64       flat over pages: thread = 12 bytes in (3 aligned u32 loads) -> 16 chars out (one uint4 store),
//                  6-bit -> ASCII by arithmetic (no table); page extents are 16-byte aligned by k_layout.
//   k_adler_seg / k_adler_fin, k_crc_flat: stage-level entry points for tests (vcp_adler32 / vcp_crc32).
#include "vcp_internal.cuh"
#include <algorithm>

namespace vcp {

namespace {

constexpr uint32_t kPoly = 0xEDB88320u;

// a(x) * b(x) mod P in the reflected representation (bit 31 = x^0), as zlib's multmodp
__device__ __forceinline__ uint32_t gf_mul(uint32_t a, uint32_t b) {
    uint32_t p = 0;
#pragma unroll 4
    for (int i = 0; i < 32; i++) {
        if (a & (0x80000000u >> i)) p ^= b;
        b = (b & 1u) ? (b >> 1) ^ kPoly : (b >> 1);
    }
    return p;
}

// x^(8*n) mod P
__device__ uint32_t gf_xpow8(unsigned long long n) {
    uint32_t r = 0x80000000u;          // x^0
    uint32_t pw = 0x00800000u;         // x^8
    while (n) {
        if (n & 1ull) r = gf_mul(pw, r);
        pw = gf_mul(pw, pw);
        n >>= 1;
    }
    return r;
}

__device__ __forceinline__ void crc_table_init(uint32_t* tab) {
    for (int n = threadIdx.x; n < 256; n += blockDim.x) {
        uint32_t c = (uint32_t)n;
#pragma unroll
        for (int k = 0; k < 8; k++) c = (c & 1u) ? (kPoly ^ (c >> 1)) : (c >> 1);
        tab[n] = c;
    }
}

// CRC-32 (zlib convention) of pre[0..npre) followed by data[0..n), computed by the whole CTA.
// Result valid in thread 0.  red: blockDim.x/32 words of shared scratch.
__device__ uint32_t cta_crc32(const uint32_t* tab, const uint8_t* pre, int npre, const uint8_t* data,
                              unsigned long long n, uint32_t* red) {
    const unsigned long long total = (unsigned long long)npre + n;
    const unsigned long long seg = (total + blockDim.x - 1) / blockDim.x;
    const unsigned long long lo = min(total, seg * threadIdx.x), hi = min(total, lo + seg);
    uint32_t c = 0;
    if (hi > lo) {
        c = 0xFFFFFFFFu;
        for (unsigned long long i = lo; i < hi; i++) {
            const uint8_t byte = i < (unsigned long long)npre ? pre[i] : data[i - npre];
            c = tab[(c ^ byte) & 0xFFu] ^ (c >> 8);
        }
        c ^= 0xFFFFFFFFu;
        const unsigned long long after = total - hi;
        if (after) c = gf_mul(gf_xpow8(after), c);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c ^= __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
    __syncthreads();
    uint32_t r = 0;
    if (threadIdx.x == 0) for (int w = 0; w < (int)(blockDim.x >> 5); w++) r ^= red[w];
    return r;
}

__device__ __forceinline__ void put_be32(uint8_t* p, uint32_t v) {
    p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v;
}

}  // namespace

// ------------------------------------------------------------------------------------------ framing
__global__ void __launch_bounds__(256) k_png_finish(BatchD B) {
    if (B.err[0]) return;
    __shared__ uint32_t tab[256];
    __shared__ uint32_t red[8];
    __shared__ uint8_t ihdr[17];
    const int b = blockIdx.x;
    const BlockD blk = B.blocks[b];
    const PageD& P = B.pages[blk.page];
    crc_table_init(tab);
    __syncthreads();
    uint8_t* pay = B.png + B.blk_dst[b];
    const uint32_t plen = B.blk_len[b];
    const uint8_t idat[4] = {'I', 'D', 'A', 'T'};
    __shared__ uint8_t s_idat[4];
    if (threadIdx.x < 4) s_idat[threadIdx.x] = idat[threadIdx.x];
    __syncthreads();
    const uint32_t crc = cta_crc32(tab, s_idat, 4, pay, plen, red);
    if (threadIdx.x == 0) {
        put_be32(pay - 8, plen);
        pay[-4] = 'I'; pay[-3] = 'D'; pay[-2] = 'A'; pay[-1] = 'T';
        put_be32(pay + plen, crc);
        if (blk.last) {
            uint8_t* e = pay + plen + 4;
            const uint8_t iend[12] = {0, 0, 0, 0, 'I', 'E', 'N', 'D', 0xAE, 0x42, 0x60, 0x82};
            for (int i = 0; i < 12; i++) e[i] = iend[i];
        }
        if (blk.first) {
            ihdr[0] = 'I'; ihdr[1] = 'H'; ihdr[2] = 'D'; ihdr[3] = 'R';
            put_be32(ihdr + 4, (uint32_t)P.w); put_be32(ihdr + 8, (uint32_t)P.h);
            ihdr[12] = 8; ihdr[13] = (uint8_t)P.color_type; ihdr[14] = 0; ihdr[15] = 0; ihdr[16] = 0;
        }
    }
    __syncthreads();
    if (blk.first) {
        const uint32_t hcrc = cta_crc32(tab, ihdr, 17, nullptr, 0, red);
        if (threadIdx.x == 0) {
            uint8_t* s = B.png + B.png_off[blk.page];
            const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
            for (int i = 0; i < 8; i++) s[i] = sig[i];
            put_be32(s + 8, 13);
            for (int i = 0; i < 17; i++) s[12 + i] = ihdr[i];
            put_be32(s + 29, hcrc);
        }
    }
}

int launch_png_finish(const BatchD& b, cudaStream_t st) {
    if (b.nblocks == 0 || !b.framed) return 0;
    k_png_finish<<<b.nblocks, 256, 0, st>>>(b);
    return 1;
}

// ------------------------------------------------------------------------------------------ base64
namespace {

__device__ __forceinline__ uint32_t b64_char(uint32_t v) {          // 6 bits -> ASCII
    // 'A'+v | 'a'+v-26 | '0'+v-52 | '+' (62) | '/' (63)
    uint32_t c = v + 65u;
    c += (v >= 26u) ? 6u : 0u;
    c -= (v >= 52u) ? 75u : 0u;
    c -= (v >= 62u) ? 15u : 0u;
    c += (v >= 63u) ? 3u : 0u;
    return c;
}

__device__ __forceinline__ uint32_t b64_quad(uint32_t b0, uint32_t b1, uint32_t b2) {   // 3 bytes -> 4 chars (LE word)
    const uint32_t t = (b0 << 16) | (b1 << 8) | b2;
    return b64_char(t >> 18) | (b64_char((t >> 12) & 63u) << 8) | (b64_char((t >> 6) & 63u) << 16) | (b64_char(t & 63u) << 24);
}

// 12 source bytes (3 LE words) -> 16 chars
__device__ __forceinline__ uint4 b64_12(uint32_t w0, uint32_t w1, uint32_t w2) {
    uint4 o;
    o.x = b64_quad(w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u);
    o.y = b64_quad(w0 >> 24, w1 & 255u, (w1 >> 8) & 255u);
    o.z = b64_quad((w1 >> 16) & 255u, w1 >> 24, w2 & 255u);
    o.w = b64_quad((w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24);
    return o;
}

// encode src[0..len) -> dst; src 4-byte aligned, dst 16-byte aligned; called with a grid-stride over units of 12 bytes
__device__ __forceinline__ void b64_unit(const uint8_t* __restrict__ src, unsigned long long len, uint8_t* __restrict__ dst,
                                         unsigned long long u) {
    const unsigned long long i = u * 12ull;
    if (i + 12ull <= len) {
        const uint32_t* s = reinterpret_cast<const uint32_t*>(src + i);
        *reinterpret_cast<uint4*>(dst + u * 16ull) = b64_12(__ldg(s), __ldg(s + 1), __ldg(s + 2));
    } else if (i < len) {                                           // ragged tail: byte-wise with '=' padding
        uint8_t* d = dst + u * 16ull;
        for (unsigned long long j = i; j < len; j += 3) {
            const uint32_t b0 = src[j], b1 = j + 1 < len ? src[j + 1] : 0u, b2 = j + 2 < len ? src[j + 2] : 0u;
            const uint32_t q = b64_quad(b0, b1, b2);
            d[0] = (uint8_t)q; d[1] = (uint8_t)(q >> 8);
            d[2] = j + 1 < len ? (uint8_t)(q >> 16) : (uint8_t)'=';
            d[3] = j + 2 < len ? (uint8_t)(q >> 24) : (uint8_t)'=';
            d += 4;
        }
    }
}

}  // namespace

__global__ void __launch_bounds__(256) k_base64_pages(BatchD B) {
    if (B.err[0]) return;
    const int p = blockIdx.y;
    const unsigned long long len = B.png_len[p];
    const uint8_t* src = B.png + B.png_off[p];
    uint8_t* dst = B.b64 + B.b64_off[p];
    const unsigned long long units = (len + 11ull) / 12ull;
    for (unsigned long long u = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; u < units;
         u += (unsigned long long)gridDim.x * blockDim.x)
        b64_unit(src, len, dst, u);
}

int launch_base64_pages(const BatchD& b, cudaStream_t st) {
    if (b.npages == 0 || !b.want_b64) return 0;
    dim3 grid(64, b.npages);
    k_base64_pages<<<grid, 256, 0, st>>>(b);
    return 1;
}

__global__ void __launch_bounds__(256) k_base64_flat(const uint8_t* __restrict__ src, unsigned long long len, uint8_t* __restrict__ dst) {
    const unsigned long long units = (len + 11ull) / 12ull;
    for (unsigned long long u = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; u < units;
         u += (unsigned long long)gridDim.x * blockDim.x)
        b64_unit(src, len, dst, u);
}

int launch_base64_flat(const uint8_t* src, uint64_t len, uint8_t* dst, cudaStream_t st) {
    if (len == 0) return 0;
    const unsigned long long units = (len + 11ull) / 12ull;
    const int blocks = (int)std::min<unsigned long long>((units + 255ull) / 256ull, 148ull * 16ull);
    k_base64_flat<<<blocks, 256, 0, st>>>(src, len, dst);
    return 1;
}

// ------------------------------------------------------------------------------------------ flat checksums (test hooks)
constexpr int kAdlerSeg = 4096;     // bytes per thread-block segment; 4096 * 255 * 4096 fits u64 comfortably

__global__ void __launch_bounds__(256) k_adler_seg(const uint8_t* __restrict__ data, unsigned long long len, uint32_t* __restrict__ part) {
    __shared__ uint32_t r1[8], r2[8];
    const unsigned long long seg0 = (unsigned long long)blockIdx.x * kAdlerSeg;
    const int n = (int)((len - seg0) < (unsigned long long)kAdlerSeg ? (len - seg0) : (unsigned long long)kAdlerSeg);
    uint32_t s1 = 0; unsigned long long s2 = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t v = data[seg0 + i];
        s1 += v; s2 += (unsigned long long)(n - i) * v;
    }
    uint32_t s2m = (uint32_t)(s2 % 65521ull);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2m += __shfl_xor_sync(0xffffffffu, s2m, o); }
    if ((threadIdx.x & 31) == 0) { r1[threadIdx.x >> 5] = s1; r2[threadIdx.x >> 5] = s2m; }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t1 = 0, t2 = 0;
        for (int w = 0; w < 8; w++) { t1 += r1[w]; t2 += r2[w]; }
        part[2 * blockIdx.x] = t1 % 65521u; part[2 * blockIdx.x + 1] = t2 % 65521u;
    }
}

__global__ void k_adler_fin(const uint32_t* __restrict__ part, unsigned long long len, uint32_t* __restrict__ out) {
    const unsigned long long nseg = (len + kAdlerSeg - 1) / kAdlerSeg;
    uint32_t a = 1, b = 0;
    for (unsigned long long k = 0; k < nseg; k++) {
        const uint32_t n = (uint32_t)((len - k * kAdlerSeg) < (unsigned long long)kAdlerSeg ? (len - k * kAdlerSeg) : (unsigned long long)kAdlerSeg);
        b = (uint32_t)((b + (unsigned long long)n * a + part[2 * k + 1]) % 65521ull);
        a = (a + part[2 * k]) % 65521u;
    }
    *out = (b << 16) | a;
}

int launch_adler_flat(const uint8_t* data, uint64_t len, uint32_t* scratch, uint32_t* out, cudaStream_t st) {
    const unsigned long long nseg = (len + kAdlerSeg - 1) / kAdlerSeg;
    int n = 0;
    if (nseg) { k_adler_seg<<<(unsigned)nseg, 256, 0, st>>>(data, len, scratch); n++; }
    k_adler_fin<<<1, 1, 0, st>>>(scratch, len, out);
    return n + 1;
}

__global__ void __launch_bounds__(256) k_crc_flat(const uint8_t* __restrict__ data, unsigned long long len, uint32_t* __restrict__ out) {
    __shared__ uint32_t tab[256];
    __shared__ uint32_t red[8];
    crc_table_init(tab);
    __syncthreads();
    const uint32_t c = cta_crc32(tab, nullptr, 0, data, len, red);
    if (threadIdx.x == 0) *out = c;
}

int launch_crc_flat(const uint8_t* data, uint64_t len, uint32_t* out, cudaStream_t st) {
    k_crc_flat<<<1, 256, 0, st>>>(data, len, out);
    return 1;
}

}  // namespace vcp
