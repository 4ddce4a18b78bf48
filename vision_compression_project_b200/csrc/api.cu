// api.cu — the C ABI of libvcprep.so (include/vcprep.h): handle, planning, arena, launch sequence.
//
// This is the boundary the reference's per-page body binds to (SURVEY.md §8 b): it stands where
//     page_image.save(page_image_path)         backend/app/pipeline/pdf_extract.py:130
//     generate_content([prompt, page_image])   backend/app/pipeline/pdf_extract.py:55  (image -> PNG blob -> base64)
// run Pillow / zlib / binascii natively.  A batch of pages goes through ONE launch set (not one per page):
//   H2D -> convert -> reduce -> resample H -> resample V -> PNG filter (+Adler partials) -> Adler combine
//       -> LZ77 -> Huffman build -> layout -> payload init -> Huffman emit -> CRC/framing -> base64 -> D2H.
// Host work is planning only: geometry, Pillow's coefficient tables (double math, cached), arena carving.
#include "../../include/vcprep.h"
#include "vcp_internal.cuh"
#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <immintrin.h>
#include <condition_variable>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <chrono>
#include <thread>
#include <tuple>
#include <vector>

using namespace vcp;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(e_ == cudaErrorMemoryAllocation ? VCP_ENOMEM : VCP_ECUDA, \
    "CUDA error %s at %s:%d (%s)", cudaGetErrorString(e_), __FILE__, __LINE__, #call); } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Coeffs { int ksize; std::vector<int32_t> bounds; std::vector<int32_t> kt; };   // kt transposed: [k][out]
typedef std::tuple<int, int, int, float, float> CoeffKey;

enum { EV_START, EV_H2D, EV_CONV, EV_PIXEL, EV_FILTER, EV_LZ, EV_HUFF, EV_ASSEMBLE, EV_B64, EV_D2H, EV_COUNT };

}  // namespace

// A lane = one CUDA stream with its own device arena and pinned staging block.  A handle has four, so that with host
// inputs the H2D copies run back to back (up to three groups ahead) while earlier groups compute (vcp_prepare_batch).
struct Lane {
    cudaStream_t stream = nullptr;
    uint8_t* arena = nullptr; size_t arena_cap = 0;
    uint8_t* meta = nullptr; size_t meta_cap = 0;          // pinned host staging: descriptors up, results down
    uint8_t* stage = nullptr; size_t stage_cap = 0;        // pinned bounce buffer for pageable page sources (filled by host threads)
    cudaEvent_t ev[EV_COUNT] = {};
};
#ifndef VCP_LANES
#define VCP_LANES 4
#endif
constexpr int kLanes = VCP_LANES;
constexpr int kMaxGroupPages = 65535;                    // the page index is grid.y / grid.z of the row kernels

struct BatchJob {                  // a streaming batch (vcp_batch_begin .. vcp_batch_end)
    struct Ready { int first, last; cudaEvent_t ev; };
    std::vector<vcp_page_desc> pages; vcp_opts opts;
    std::thread th;
    std::mutex m; std::condition_variable cv;
    std::vector<Ready> ready; size_t next = 0;
    bool done = false; int rc = 0; std::string err;
};

struct vcp_handle {
    int device = 0;
    std::mutex mu;
    BatchJob* job = nullptr;
    Lane lane[kLanes];
    std::map<CoeffKey, std::shared_ptr<const Coeffs>> coeff_cache;   // plans hold shared_ptrs: eviction never frees tables a group still uses
    vcp_stats stats = {};
    size_t group_bytes = (size_t)2 << 30;                  // max filtered bytes per launch set
    cudaEvent_t trace_t0 = nullptr;                        // VCP_TRACE: recorded at the start of a call
    size_t pipe_bytes = (size_t)96 << 20;                  // host inputs: source bytes per pipelined group
    int copy_threads = 8;                                  // host threads that fill the bounce buffer
};

namespace {

// Entry points serialise on the handle's mutex.  While a streaming batch (vcp_batch_begin .. vcp_batch_end) is in flight the
// handle belongs to its worker thread: every other call is refused with VCP_EINVAL instead of blocking (the mutex is not held
// across begin .. end, so a second call from the same thread cannot deadlock).
struct HandleGuard {
    vcp_handle* h; bool ok;
    explicit HandleGuard(vcp_handle* h_) : h(h_) { h->mu.lock(); ok = h->job == nullptr; if (!ok) h->mu.unlock(); }
    ~HandleGuard() { if (ok) h->mu.unlock(); }
    HandleGuard(const HandleGuard&) = delete;
};
#define LOCK_HANDLE(h) HandleGuard guard_(h); \
    if (!guard_.ok) return fail(VCP_EINVAL, "a streaming batch is in flight on this handle (finish it with vcp_batch_end first)")

// NVTX ranges around the host-side phases of a batch (issue of a launch set, wait + hand-over, PNG decode), for nsys / Nsight timelines.
// Header-only NVTX3: without a profiler attached a range is a call through a null function table.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
};

int ensure_arena(Lane& L, size_t need) {
    if (need <= L.arena_cap) return 0;
    if (L.arena) { CU(cudaStreamSynchronize(L.stream)); CU(cudaFree(L.arena)); L.arena = nullptr; L.arena_cap = 0; }
    const size_t cap = align_up(need + need / 8, (size_t)1 << 20);
    CU(cudaMalloc(&L.arena, cap));
    L.arena_cap = cap;
    return 0;
}

int ensure_meta(Lane& L, size_t need) {
    if (need <= L.meta_cap) return 0;
    if (L.meta) { CU(cudaStreamSynchronize(L.stream)); CU(cudaFreeHost(L.meta)); L.meta = nullptr; L.meta_cap = 0; }
    const size_t cap = align_up(need * 2, (size_t)1 << 16);
    CU(cudaMallocHost(&L.meta, cap));
    L.meta_cap = cap;
    return 0;
}

int ensure_stage(Lane& L, size_t need) {
    if (need <= L.stage_cap) return 0;
    if (L.stage) { CU(cudaStreamSynchronize(L.stream)); CU(cudaFreeHost(L.stage)); L.stage = nullptr; L.stage_cap = 0; }
    const size_t cap = align_up(need + need / 4, (size_t)1 << 20);
    CU(cudaMallocHost(&L.stage, cap));
    L.stage_cap = cap;
    return 0;
}

// ---- host worker pool.  The byte work the host does per launch set (packing pageable page sources into the pinned bounce buffer,
// filling the caller's bytes objects) runs on a few persistent threads: spawning 16 threads per 96 MB group cost ~0.5 ms each time.
// One job at a time (callers queue on job_m); the calling thread works too.  Leaked on purpose: no destructor order at exit.
class HostPool {
public:
    static HostPool& get() { static HostPool* p = new HostPool(); return *p; }
    // runs fn(0) .. fn(parts - 1), at most `threads` of them at once
    void run(int parts, int threads, const std::function<void(int)>& fn) {
        if (parts <= 0) return;
        threads = std::max(1, std::min({threads, parts, kMaxThreads + 1}));
        if (threads == 1) { for (int i = 0; i < parts; i++) fn(i); return; }
        std::lock_guard<std::mutex> job(job_m);
        grow(threads - 1);
        {
            std::lock_guard<std::mutex> g(m);
            cur = &fn; total = parts; next = 0; pending = parts; helpers = threads - 1; epoch++;
        }
        cv_work.notify_all();
        work();
        std::unique_lock<std::mutex> g(m);
        cv_done.wait(g, [&] { return pending == 0; });
        cur = nullptr;
    }

private:
    static constexpr int kMaxThreads = 32;
    std::mutex job_m, m;
    std::condition_variable cv_work, cv_done;
    const std::function<void(int)>* cur = nullptr;
    int total = 0, next = 0, pending = 0, helpers = 0, nthreads = 0;
    unsigned long long epoch = 0;

    void grow(int want) {
        while (nthreads < std::min(want, kMaxThreads)) {
            const int id = nthreads++;
            std::thread([this, id] { loop(id); }).detach();
        }
    }
    void work() {
        for (;;) {
            int i;
            const std::function<void(int)>* f;
            { std::lock_guard<std::mutex> g(m); if (!cur || next >= total) return; i = next++; f = cur; }
            (*f)(i);
            { std::lock_guard<std::mutex> g(m); if (--pending == 0) cv_done.notify_all(); }
        }
    }
    void loop(int id) {
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> g(m);
                cv_work.wait(g, [&] { return epoch != seen; });
                seen = epoch;
                if (id >= helpers) continue;               // this job asked for fewer threads
            }
            work();
        }
    }
};

// copy n byte ranges on T host threads, splitting the total evenly (a page of rows counts as one range per row when strided).
// drop4 ranges are RGBX / RGBA pixels (Pillow's 4-byte storage of RGB images) of which only the first three bytes of every pixel are
// wanted: len counts SOURCE bytes (a multiple of 4), the destination receives 3/4 of them.
struct CopyJob { uint8_t* dst; const uint8_t* src; size_t len; bool drop4; };

__attribute__((target("ssse3"))) void pack_rgbx_ssse3(uint8_t* d, const uint8_t* s, size_t npix) {
    const __m128i sh = _mm_setr_epi8(0, 1, 2, 4, 5, 6, 8, 9, 10, 12, 13, 14, -1, -1, -1, -1);
    size_t i = 0;
    const bool aligned = ((uintptr_t)d & 15) == 0;
    for (; i + 16 <= npix; i += 16) {                         // 64 source bytes -> 48
        const __m128i a = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i*)(s + 4 * i)), sh);
        const __m128i b = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i*)(s + 4 * i + 16)), sh);
        const __m128i c = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i*)(s + 4 * i + 32)), sh);
        const __m128i e = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i*)(s + 4 * i + 48)), sh);
        const __m128i o0 = _mm_or_si128(a, _mm_slli_si128(b, 12));
        const __m128i o1 = _mm_or_si128(_mm_srli_si128(b, 4), _mm_slli_si128(c, 8));
        const __m128i o2 = _mm_or_si128(_mm_srli_si128(c, 8), _mm_slli_si128(e, 4));
        __m128i* o = (__m128i*)(d + 3 * i);
        if (aligned) { _mm_stream_si128(o, o0); _mm_stream_si128(o + 1, o1); _mm_stream_si128(o + 2, o2); }   // the bounce buffer is read next by the DMA engine, not by this core
        else { _mm_storeu_si128(o, o0); _mm_storeu_si128(o + 1, o1); _mm_storeu_si128(o + 2, o2); }
    }
    for (; i < npix; i++) { d[3 * i] = s[4 * i]; d[3 * i + 1] = s[4 * i + 1]; d[3 * i + 2] = s[4 * i + 2]; }
    if (aligned) _mm_sfence();
}

// memcpy into the pinned bounce buffer with streaming stores.  The copy engine reads that buffer next: lines left dirty in the CPU
// caches by ordinary stores make its reads snoop them (measured on the B200 box: a one-page launch set took 1.6 ms longer from a
// pageable numpy array than from a PIL image, whose pack already streamed), and the data would only evict useful cache lines.
__attribute__((target("sse2"))) void stream_copy(uint8_t* d, const uint8_t* s, size_t n) {
    if (n < 256) { memcpy(d, s, n); return; }
    const size_t head = (16 - ((uintptr_t)d & 15)) & 15;
    if (head) { memcpy(d, s, head); d += head; s += head; n -= head; }
    size_t i = 0;
    for (; i + 64 <= n; i += 64) {
        const __m128i a = _mm_loadu_si128((const __m128i*)(s + i)), b = _mm_loadu_si128((const __m128i*)(s + i + 16));
        const __m128i c = _mm_loadu_si128((const __m128i*)(s + i + 32)), e = _mm_loadu_si128((const __m128i*)(s + i + 48));
        _mm_stream_si128((__m128i*)(d + i), a); _mm_stream_si128((__m128i*)(d + i + 16), b);
        _mm_stream_si128((__m128i*)(d + i + 32), c); _mm_stream_si128((__m128i*)(d + i + 48), e);
    }
    _mm_sfence();
    if (i < n) memcpy(d + i, s + i, n - i);
}

void pack_rgbx(uint8_t* d, const uint8_t* s, size_t npix) {
    static const bool has = __builtin_cpu_supports("ssse3");
    if (has) { pack_rgbx_ssse3(d, s, npix); return; }
    for (size_t i = 0; i < npix; i++) { d[3 * i] = s[4 * i]; d[3 * i + 1] = s[4 * i + 1]; d[3 * i + 2] = s[4 * i + 2]; }
}

void parallel_copy(const std::vector<CopyJob>& jobs, int T) {
    size_t total = 0;
    for (auto& j : jobs) total += j.len;
    T = std::max(1, std::min(T, 16));
    if (total < ((size_t)1 << 20)) T = 1;
    // part t takes the source bytes [t * total / T, (t + 1) * total / T) of the concatenation, cut on 64-byte (16-pixel) boundaries
    // of each range so that packed pixels never straddle two parts
    std::vector<size_t> cum(jobs.size() + 1, 0);              // many small ranges (one per row of a strided source): every part finds its first range by bisection
    for (size_t i = 0; i < jobs.size(); i++) cum[i + 1] = cum[i] + jobs[i].len;
    HostPool::get().run(T, T, [&jobs, &cum, total, T](int t) {
        const size_t lo = total * t / T, hi = total * (t + 1) / T;
        size_t i0 = (size_t)(std::upper_bound(cum.begin(), cum.end(), lo) - cum.begin());
        i0 = i0 ? i0 - 1 : 0;
        size_t pos = cum[i0];
        for (size_t i = i0; i < jobs.size() && pos < hi; i++) {
            const CopyJob& j = jobs[i];
            size_t a = std::max(lo, pos) - pos, b = std::min(hi, pos + j.len) - pos;
            if (std::max(lo, pos) < std::min(hi, pos + j.len)) {
                if (a) a = (a + 63) & ~(size_t)63;
                if (b < j.len) b = (b + 63) & ~(size_t)63;
                a = std::min(a, j.len); b = std::min(b, j.len);
                if (a < b) {
                    if (j.drop4) pack_rgbx(j.dst + a / 4 * 3, j.src + a, (b - a) / 4);
                    else stream_copy(j.dst + a, j.src + a, b - a);
                }
            }
            pos += j.len;
        }
    });
}

typedef std::shared_ptr<const Coeffs> CoeffRef;

CoeffRef get_coeffs(vcp_handle* h, int in_size, int out_size, int filter, float b0, float b1) {
    CoeffKey key(in_size, out_size, filter, b0, b1);
    auto it = h->coeff_cache.find(key);
    if (it != h->coeff_cache.end()) return it->second;
    if (h->coeff_cache.size() > 256) h->coeff_cache.clear();     // page plans of the group being planned keep their own references
    auto c = std::make_shared<Coeffs>();
    if (resample_coeffs_host(in_size, out_size, filter, b0, b1, nullptr, nullptr, &c->ksize) != 0) return nullptr;
    std::vector<int32_t> kk((size_t)out_size * c->ksize);
    c->bounds.resize((size_t)out_size * 2);
    if (resample_coeffs_host(in_size, out_size, filter, b0, b1, c->bounds.data(), kk.data(), &c->ksize) != 0) return nullptr;
    c->kt.resize(kk.size());
    for (int x = 0; x < out_size; x++)
        for (int k = 0; k < c->ksize; k++) c->kt[(size_t)k * out_size + x] = kk[(size_t)x * c->ksize + k];
    return h->coeff_cache[key] = c;
}

struct Bump {
    size_t off = 0;
    size_t take(size_t n, size_t a = 256) { off = align_up(off, a); const size_t r = off; off += n; return r; }
};

constexpr size_t kNone = (size_t)-1;

// Host-side plan of one page (offsets into the arena; kNone = stage skipped).
struct PagePlan {
    int status = 0;
    const uint8_t* src = nullptr; int64_t src_stride = 0;
    const uint8_t* const* row_ptrs = nullptr;      // host source addressed through a row table (Pillow's multi-block images)
    int sw = 0, sh = 0, sc = 0, c = 0, pc = 0, fx = 1, fy = 1, rw = 0, rh = 0, w = 0, h = 0;
    bool need_conv = false, need_red = false, need_h = false, need_v = false;
    bool staged = false;     // host source in pageable memory: packed into the lane's pinned bounce buffer by host threads
    int dsc = 0;             // channels of the page as it arrives on the device (3 when the host pack dropped the X of RGBX storage)
    float box[4] = {0, 0, 0, 0};
    size_t o_raw = kNone, o_conv = kNone, o_red = kNone, o_tmp = kNone, o_vout = kNone, o_filt = kNone;
    size_t o_hb = kNone, o_hk = kNone, o_vb = kNone, o_vk = kNone;
    CoeffRef ch, cv;
    int64_t filt_len = 0;
    int nblk = 0, nsub = 0;
    uint64_t png_bound = 0;
};

uint64_t png_bound_of(int64_t filt_len, int nblk) {
    return align_up((size_t)(8 + 25 + 12 + (uint64_t)nblk * (12 + 6) + filt_len + 5 * (uint64_t)(filt_len / 65535 + nblk + 1)), 16);
}
uint64_t b64_bound_of(uint64_t png_bound) { return align_up((size_t)(4 * ((png_bound + 2) / 3)), 16); }

int color_type_of(int c) { return c == 1 ? 0 : c == 2 ? 4 : c == 3 ? 2 : 6; }

// geometry + validation of one page; no allocation
int plan_geometry(const vcp_page_desc& d, const vcp_opts& o, PagePlan& P) {
    if (!d.src && !d.row_ptrs) return fail(VCP_EINVAL, "page src is NULL");
    if (d.row_ptrs && o.src_device) return fail(VCP_EINVAL, "row_ptrs describes host memory; device pages are strided");
    if (d.width <= 0 || d.height <= 0) return fail(VCP_EINVAL, "bad page size %dx%d", d.width, d.height);
    if (d.channels < 1 || d.channels > 4) return fail(VCP_EINVAL, "unsupported channel count %d (1=L 2=LA 3=RGB 4=RGBA)", d.channels);
    if ((int64_t)d.width * d.channels > (1 << 20)) return fail(VCP_EINVAL, "page too wide (%d px)", d.width);
    if (d.height > (1 << 20)) return fail(VCP_EINVAL, "page too tall (%d px)", d.height);
    P.src = (const uint8_t*)d.src; P.sw = d.width; P.sh = d.height; P.sc = d.channels;
    P.row_ptrs = (const uint8_t* const*)d.row_ptrs;
    if (P.row_ptrs) P.src = P.row_ptrs[0];
    P.src_stride = d.row_stride ? d.row_stride : (int64_t)d.width * d.channels;
    if (P.src_stride < (int64_t)d.width * d.channels) return fail(VCP_EINVAL, "row_stride %lld smaller than a row", (long long)d.row_stride);
    P.c = o.out_channels ? o.out_channels : d.channels;
    if (P.c != 1 && P.c != 3 && P.c != d.channels) return fail(VCP_EINVAL, "unsupported output channel count %d", P.c);
    P.pc = (P.sc <= 2 && P.c == 3) ? 1 : P.c;        // gray sources stay single-channel until the PNG filter replicates them
    P.need_conv = P.pc != P.sc;
    P.fx = d.reduce_x > 1 ? d.reduce_x : 1; P.fy = d.reduce_y > 1 ? d.reduce_y : 1;
    P.need_red = P.fx > 1 || P.fy > 1;
    P.rw = (P.sw + P.fx - 1) / P.fx; P.rh = (P.sh + P.fy - 1) / P.fy;
    const bool resize = d.dst_width > 0 && d.dst_height > 0;
    if ((d.dst_width > 0) != (d.dst_height > 0)) return fail(VCP_EINVAL, "dst_width/dst_height must both be set or both be 0");
    P.w = resize ? d.dst_width : P.rw; P.h = resize ? d.dst_height : P.rh;
    if ((int64_t)P.w * P.c > 65536 || P.h > 65535 || P.sh > 65535) return fail(VCP_EINVAL, "page too large for the row kernels (%dx%d -> %dx%d)", P.sw, P.sh, P.w, P.h);
    P.box[0] = 0.f; P.box[1] = 0.f;
    P.box[2] = P.need_red ? (float)((double)P.sw / P.fx) : (float)P.rw;
    P.box[3] = P.need_red ? (float)((double)P.sh / P.fy) : (float)P.rh;
    P.need_h = P.w != P.rw || P.box[2] != (float)P.w;
    P.need_v = P.h != P.rh || P.box[3] != (float)P.h;
    if ((P.need_h || P.need_v) && (P.c == 2 || P.c == 4))
        return fail(VCP_EINVAL, "resize of images with alpha is not on the path (Pillow premultiplies); convert to RGB or L");
    if ((P.need_h || P.need_v) && (o.resample < VCP_LANCZOS || o.resample > VCP_HAMMING))
        return fail(VCP_EINVAL, "unsupported resample filter %d", o.resample);
    P.filt_len = (int64_t)P.h * (1 + (int64_t)P.w * P.c);
    if (P.filt_len >= ((int64_t)1 << 31) - (1 << 20)) return fail(VCP_EINVAL, "page too large (%lld filtered bytes)", (long long)P.filt_len);
    P.nblk = (int)((P.filt_len + kBlockBytes - 1) / kBlockBytes);
    P.nsub = 0;
    for (int b = 0; b < P.nblk; b++) {
        const int64_t len = std::min<int64_t>(kBlockBytes, P.filt_len - (int64_t)b * kBlockBytes);
        P.nsub += (int)((len + kSubBytes - 1) / kSubBytes);
    }
    P.png_bound = png_bound_of(P.filt_len, P.nblk);
    return 0;
}

struct GroupOut {       // where the launch set of a group left its results (pinned host copies)
    const uint64_t* png_off; const uint64_t* png_len; const uint64_t* b64_off; const uint64_t* b64_len;
    const uint32_t* adler; const uint64_t* totals;
    uint8_t* d_png; uint8_t* d_b64; uint8_t* d_filt0; uint32_t* d_tokens; uint32_t* d_sub_ntok; uint32_t* d_sub_hist;
    int nsub; int nblocks;
    std::vector<size_t> filt_off;     // per page: offset of its filtered stream from d_filt0's arena base
    // issue -> finish state
    Lane* lane = nullptr; int n = 0; size_t res_bytes = 0; uint8_t* host_res = nullptr;
    uint64_t launches = 0, in_bytes = 0, filt_bytes = 0;
};

// Runs the whole launch set for pages[0..n) (all with status 0).  framed=0 + stream input: the pages' filtered
// streams are given directly (stage-level vcp_deflate / vcp_lz_tokens): plans[i].src is the stream, filt_len set.
enum RunMode { RUN_FULL = 0, RUN_STREAM = 1, RUN_FILTER_ONLY = 2, RUN_LZ_ONLY = 3 };

// Enqueues the whole launch set of a group on lane L (nothing here waits for the GPU).
int issue_group(vcp_handle* h, Lane& L, std::vector<PagePlan>& plans, const vcp_opts& o, RunMode mode, GroupOut& out) {
    NvtxRange nvtx_("vcp issue launch set");
    const int n = (int)plans.size();
    cudaStream_t st = L.stream;
    const bool stream_in = (mode == RUN_STREAM || mode == RUN_LZ_ONLY);
    // ---------------- carve the arena
    Bump bump;
    int nblocks = 0, nsub = 0, nrows = 0;
    uint64_t png_cap = 0, b64_cap = 0;
    std::vector<int32_t> coeff_blob;
    auto put_coeff = [&](const std::vector<int32_t>& v) { const size_t o2 = coeff_blob.size(); coeff_blob.insert(coeff_blob.end(), v.begin(), v.end()); return o2; };
    std::map<const Coeffs*, std::pair<size_t, size_t>> coeff_at;   // -> (bounds idx, kt idx) in coeff_blob; the plans' references keep the keys alive
    for (auto& P : plans) {
        P.dsc = P.sc;
        if (!stream_in && !o.src_device) {
            // Pinned (or registered) sources are DMA'd where they lie.  Pageable ones — e.g. Pillow's own pixel storage — would make
            // cudaMemcpyAsync stage them on one thread at ~11 GB/s: host threads pack them into the lane's pinned bounce buffer instead
            // (the previous group's DMA and kernels run meanwhile).  Pillow keeps RGB images as 4-byte RGBX pixels: the pack drops
            // the X there and then (3/4 of the bytes cross PCIe, and the convert kernel has nothing left to do).
            cudaPointerAttributes at;
            const bool pinned = !P.row_ptrs && cudaPointerGetAttributes(&at, P.src) == cudaSuccess &&
                                (at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged);
            cudaGetLastError();
            P.staged = !pinned;
            if (P.staged && P.sc == 4 && P.c == 3) { P.dsc = 3; P.need_conv = false; }
        }
        if (!stream_in) {
            if (!o.src_device) P.o_raw = bump.take((size_t)P.sw * P.dsc * P.sh + 16);
            if (P.need_conv) P.o_conv = bump.take((size_t)P.sw * P.sh * P.pc + 16);
            if (P.need_red) P.o_red = bump.take((size_t)P.rw * P.rh * P.pc + 16);
            if (P.need_h) P.o_tmp = bump.take(align_up((size_t)P.w * P.pc, 16) * P.rh + 16);      // rows padded to 16 B: aligned word loads in the V pass
            if (P.need_v) P.o_vout = bump.take(align_up((size_t)P.w * P.pc, 16) * P.h + 16);
            if (P.need_h) {
                P.ch = get_coeffs(h, P.rw, P.w, o.resample, P.box[0], P.box[2]);
                if (!P.ch) return fail(VCP_EINVAL, "bad resample parameters");
                if (!coeff_at.count(P.ch.get())) { const size_t a = put_coeff(P.ch->bounds); const size_t b = put_coeff(P.ch->kt); coeff_at[P.ch.get()] = {a, b}; }
            }
            if (P.need_v) {
                P.cv = get_coeffs(h, P.rh, P.h, o.resample, P.box[1], P.box[3]);
                if (!P.cv) return fail(VCP_EINVAL, "bad resample parameters");
                if (!coeff_at.count(P.cv.get())) { const size_t a = put_coeff(P.cv->bounds); const size_t b = put_coeff(P.cv->kt); coeff_at[P.cv.get()] = {a, b}; }
            }
        }
        nblocks += P.nblk; nsub += P.nsub; nrows += P.h;
        png_cap += P.png_bound; b64_cap += b64_bound_of(P.png_bound);
    }
    // filtered streams: one contiguous region so that token index = byte offset from its base
    const size_t o_filt_region = bump.take(0);
    for (auto& P : plans) { bump.take(kStreamPad); P.o_filt = bump.take((size_t)P.filt_len); }
    bump.take(kStreamPad + 2048);                 // run_end() reads up to 1 KiB + a window past the last stream
    const size_t filt_region_bytes = align_up(bump.off - o_filt_region, 256);
    const bool need_lz = mode != RUN_FILTER_ONLY;
    const bool need_huff = mode == RUN_FULL || mode == RUN_STREAM;
    const size_t o_tokens = need_lz ? bump.take(filt_region_bytes * 4) : kNone;
    const size_t o_sub_ntok = bump.take((size_t)nsub * 4 + 4);
    const size_t o_sub_hist = bump.take((size_t)nsub * kHistSize * 4 + 4);
    const size_t o_row_adler = bump.take((size_t)nrows * 4 + 4);
    const size_t o_row_busy = bump.take((size_t)nrows + 4);
    const size_t o_lz_order = bump.take((size_t)nsub * 5 + 32 + ((size_t)nsub / 256 + 2) * 64 * 4);   // u32 order + u8 keys + per-CTA key counts of the stable sort
    const size_t o_page_adler = bump.take((size_t)n * 4);
    const size_t o_blk_code = bump.take((size_t)nblocks * kCodeStride * 2 + 4);
    const size_t o_blk_clen = bump.take((size_t)nblocks * kCodeStride + 4);
    const size_t o_blk_hdr = bump.take((size_t)nblocks * kHdrBytes + 4);
    const size_t o_blk_hdr_bits = bump.take((size_t)nblocks * 4 + 4);
    const size_t o_blk_eob = bump.take((size_t)nblocks * 8 + 8);
    const size_t o_blk_bits = bump.take((size_t)nblocks * 8 + 8);
    const size_t o_blk_stored = bump.take((size_t)nblocks * 4 + 4);
    const size_t o_blk_len = bump.take((size_t)nblocks * 4 + 4);
    const size_t o_sub_bitoff = bump.take((size_t)nsub * 8 + 8);
    const size_t o_blk_dst = bump.take((size_t)nblocks * 8 + 8);
    // results block (copied down in one piece): png_off[n] png_len[n] b64_off[n] b64_len[n] totals[2] adler[n] err[2]
    const size_t res_bytes = (size_t)n * 8 * 4 + 16 + align_up((size_t)n * 4, 8) + 8;
    const size_t o_res = bump.take(res_bytes);
    const size_t o_cnt = bump.take(1024);
    const size_t o_png = need_huff ? bump.take((size_t)png_cap + 64) : kNone;
    const size_t o_b64 = (need_huff && o.want_b64) ? bump.take((size_t)b64_cap + 64) : kNone;
    const size_t o_coeff = bump.take(coeff_blob.size() * 4 + 4);
    const size_t o_pages = bump.take((size_t)n * sizeof(PageD));
    const size_t o_blocks = bump.take((size_t)nblocks * sizeof(BlockD) + 8);
    const size_t o_sub2blk = bump.take((size_t)nsub * 4 + 4);
    const size_t o_item2sub = bump.take((size_t)nsub * 4 + 4);
    const size_t desc_bytes = bump.off - o_coeff;
    int rc = ensure_arena(L, bump.off + 256);
    if (rc) return rc;
    rc = ensure_meta(L, desc_bytes + res_bytes + 256);
    if (rc) return rc;
    uint8_t* A = L.arena;
    h->stats.arena_bytes = 0; for (const Lane& l : h->lane) h->stats.arena_bytes += l.arena_cap;

    // ---------------- descriptors (built in pinned memory, mirrored layout of [o_coeff, end))
    uint8_t* M = L.meta;
    memset(M, 0, desc_bytes);
    if (!coeff_blob.empty()) memcpy(M, coeff_blob.data(), coeff_blob.size() * 4);
    PageD* hp = reinterpret_cast<PageD*>(M + (o_pages - o_coeff));
    BlockD* hb = reinterpret_cast<BlockD*>(M + (o_blocks - o_coeff));
    uint32_t* hs2b = reinterpret_cast<uint32_t*>(M + (o_sub2blk - o_coeff));
    uint32_t* hi2s = reinterpret_cast<uint32_t*>(M + (o_item2sub - o_coeff));
    int nitems = 0;
    const int32_t* d_coeff = reinterpret_cast<const int32_t*>(A + o_coeff);
    int blk = 0, sub = 0, row = 0;
    int max_sh = 0, max_sw = 0, max_rh = 0, max_rw = 0, max_w = 0, max_h = 0, max_wc = 0;
    bool any_conv = false, any_red = false, any_h = false, any_v = false;
    uint64_t in_bytes = 0, filt_bytes = 0;
    out.filt_off.clear();
    for (int i = 0; i < n; i++) {
        PagePlan& P = plans[i];
        PageD& D = hp[i];
        D.sw = P.sw; D.sh = P.sh; D.sc = P.dsc; D.c = P.c; D.pc = P.pc; D.fx = P.fx; D.fy = P.fy; D.rw = P.rw; D.rh = P.rh; D.w = P.w; D.h = P.h;
        D.color_type = color_type_of(P.c);
        const uint8_t* cur = nullptr; int64_t cur_stride = 0;
        if (!stream_in) {
            if (o.src_device) { cur = P.src; cur_stride = P.src_stride; }
            else { cur = A + P.o_raw; cur_stride = (int64_t)P.sw * P.dsc; }
            D.src = cur; D.src_stride = cur_stride;
            if (P.need_conv) { D.conv = A + P.o_conv; cur = D.conv; cur_stride = (int64_t)P.sw * P.pc; any_conv = true; }
            D.rdin = cur; D.rdin_stride = cur_stride;
            if (P.need_red) { D.red = A + P.o_red; cur = D.red; cur_stride = (int64_t)P.rw * P.pc; any_red = true; }
            D.hin = cur; D.hin_stride = cur_stride;
            if (P.need_h) {
                D.tmp = A + P.o_tmp; D.tmp_stride = (int64_t)align_up((size_t)P.w * P.pc, 16); cur = D.tmp; cur_stride = D.tmp_stride; any_h = true;
                D.hb = d_coeff + coeff_at[P.ch.get()].first; D.hk = d_coeff + coeff_at[P.ch.get()].second; D.hks = P.ch->ksize;
            }
            D.vin = cur; D.vin_stride = cur_stride;
            if (P.need_v) {
                D.vout = A + P.o_vout; D.vout_stride = (int64_t)align_up((size_t)P.w * P.pc, 16); cur = D.vout; cur_stride = D.vout_stride; any_v = true;
                D.vb = d_coeff + coeff_at[P.cv.get()].first; D.vk = d_coeff + coeff_at[P.cv.get()].second; D.vks = P.cv->ksize;
            }
            D.pix = cur; D.pix_stride = cur_stride;
            max_sh = std::max(max_sh, P.sh); max_sw = std::max(max_sw, P.sw);
            if (P.need_red) { max_rh = std::max(max_rh, P.rh); max_rw = std::max(max_rw, P.rw); }
            in_bytes += (uint64_t)P.sw * P.sh * P.sc;
        }
        max_h = std::max(max_h, P.h); max_w = std::max(max_w, P.w); max_wc = std::max(max_wc, P.w * P.c);
        D.filt = A + P.o_filt; D.filt_len = P.filt_len;
        out.filt_off.push_back(P.o_filt);
        D.row0 = row; row += P.h;
        D.blk0 = blk; D.nblk = P.nblk;
        for (int k = 0; k < P.nsub; k += kGroupSubs) hi2s[nitems++] = (uint32_t)(sub + k);   // groups restart at every page
        for (int b = 0; b < P.nblk; b++) {
            BlockD& Bk = hb[blk];
            Bk.page = i; Bk.first = b == 0; Bk.last = b == P.nblk - 1;
            Bk.start = (int64_t)b * kBlockBytes; Bk.len = std::min<int64_t>(kBlockBytes, P.filt_len - Bk.start);
            Bk.sub0 = sub; Bk.nsub = (int)((Bk.len + kSubBytes - 1) / kSubBytes);
            for (int s = 0; s < Bk.nsub; s++) hs2b[sub++] = (uint32_t)blk;
            blk++;
        }
        filt_bytes += (uint64_t)P.filt_len;
    }
    int max_resh_rows = 0;      // rows the horizontal pass walks = rh of pages that need it
    for (auto& P : plans) if (P.need_h) max_resh_rows = std::max(max_resh_rows, P.rh);

    BatchD B = {};
    B.pages = reinterpret_cast<const PageD*>(A + o_pages); B.npages = n;
    B.blocks = reinterpret_cast<const BlockD*>(A + o_blocks); B.nblocks = nblocks;
    B.sub2blk = reinterpret_cast<const uint32_t*>(A + o_sub2blk); B.nsub = nsub;
    B.item2sub = reinterpret_cast<const uint32_t*>(A + o_item2sub); B.nitems = nitems;
    B.filt_base = A + o_filt_region;
    B.tokens = need_lz ? reinterpret_cast<uint32_t*>(A + o_tokens) : nullptr;
    B.sub_ntok = reinterpret_cast<uint32_t*>(A + o_sub_ntok);
    B.sub_hist = reinterpret_cast<uint32_t*>(A + o_sub_hist);
    B.row_adler = reinterpret_cast<uint32_t*>(A + o_row_adler);
    B.row_busy = stream_in ? nullptr : A + o_row_busy;
    B.lz_order = (stream_in || nsub < 4096) ? nullptr : reinterpret_cast<uint32_t*>(A + o_lz_order);   // small sets: order does not matter
    B.blk_code = reinterpret_cast<uint16_t*>(A + o_blk_code);
    B.blk_clen = A + o_blk_clen;
    B.blk_hdr = A + o_blk_hdr;
    B.blk_hdr_bits = reinterpret_cast<uint32_t*>(A + o_blk_hdr_bits);
    B.blk_eob_bit = reinterpret_cast<uint64_t*>(A + o_blk_eob);
    B.blk_body_bits = reinterpret_cast<uint64_t*>(A + o_blk_bits);
    B.blk_stored = reinterpret_cast<uint32_t*>(A + o_blk_stored);
    B.blk_len = reinterpret_cast<uint32_t*>(A + o_blk_len);
    B.sub_bitoff = reinterpret_cast<uint64_t*>(A + o_sub_bitoff);
    B.blk_dst = reinterpret_cast<uint64_t*>(A + o_blk_dst);
    uint8_t* R = A + o_res;
    B.png_off = reinterpret_cast<uint64_t*>(R); B.png_len = B.png_off + n; B.b64_off = B.png_len + n; B.b64_len = B.b64_off + n;
    B.totals = B.b64_len + n;
    B.page_adler = reinterpret_cast<uint32_t*>(B.totals + 2);
    B.err = reinterpret_cast<uint32_t*>(R + res_bytes - 8);
    B.counters = reinterpret_cast<uint32_t*>(A + o_cnt);
    (void)o_page_adler;
    B.png = need_huff ? A + o_png : nullptr; B.png_cap = png_cap;
    B.b64 = (need_huff && o.want_b64) ? A + o_b64 : nullptr; B.b64_cap = b64_cap;
    B.framed = mode == RUN_FULL; B.level = o.compress_level; B.want_b64 = (need_huff && o.want_b64) ? 1 : 0;

    // ---------------- launch set
    uint64_t launches = 0;
    CU(cudaEventRecord(L.ev[EV_START], st));
    CU(cudaMemcpyAsync(A + o_coeff, M, desc_bytes, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(R, 0, (o_cnt + 1024) - o_res, st));
    if (!stream_in && !o.src_device) {
        std::vector<CopyJob> jobs;
        std::vector<size_t> stage_off(n, kNone);
        size_t stage_need = 0;
        for (int i = 0; i < n; i++)
            if (plans[i].staged) { stage_off[i] = stage_need; stage_need += align_up((size_t)plans[i].sw * plans[i].dsc * plans[i].sh, 256); }
        if (stage_need) {
            rc = ensure_stage(L, stage_need);
            if (rc) return rc;
            for (int i = 0; i < n; i++) {
                if (stage_off[i] == kNone) continue;
                const PagePlan& P = plans[i];
                const size_t srow = (size_t)P.sw * P.sc, drow = (size_t)P.sw * P.dsc;
                const bool drop4 = P.dsc != P.sc;
                uint8_t* d = L.stage + stage_off[i];
                if (P.row_ptrs) {
                    // a row table: consecutive rows that happen to lie back to back (all rows of one Pillow block) go as one range
                    for (int y = 0; y < P.sh;) {
                        int y1 = y + 1;
                        while (y1 < P.sh && P.row_ptrs[y1] == P.row_ptrs[y1 - 1] + srow) y1++;
                        jobs.push_back({d + (size_t)y * drow, P.row_ptrs[y], srow * (size_t)(y1 - y), drop4});
                        y = y1;
                    }
                } else if ((size_t)P.src_stride == srow) jobs.push_back({d, P.src, srow * P.sh, drop4});
                else for (int y = 0; y < P.sh; y++) jobs.push_back({d + (size_t)y * drow, P.src + (size_t)y * P.src_stride, srow, drop4});
            }
            parallel_copy(jobs, h->copy_threads);
        }
        for (int i = 0; i < n; i++) {
            const PagePlan& P = plans[i];
            const size_t rowb = (size_t)P.sw * P.dsc;
            if (stage_off[i] != kNone) CU(cudaMemcpyAsync(A + P.o_raw, L.stage + stage_off[i], rowb * P.sh, cudaMemcpyHostToDevice, st));
            else if ((size_t)P.src_stride == rowb) CU(cudaMemcpyAsync(A + P.o_raw, P.src, rowb * P.sh, cudaMemcpyHostToDevice, st));
            else CU(cudaMemcpy2DAsync(A + P.o_raw, rowb, P.src, (size_t)P.src_stride, rowb, (size_t)P.sh, cudaMemcpyHostToDevice, st));
        }
    }
    if (stream_in) {
        for (auto& P : plans)
            CU(cudaMemcpyAsync(A + P.o_filt, P.src, (size_t)P.filt_len, cudaMemcpyDeviceToDevice, st));
    }
    CU(cudaEventRecord(L.ev[EV_H2D], st));
    if (!stream_in) {
        if (any_conv) launches += launch_convert(B.pages, n, max_sh, max_sw, st);
        if (any_red) launches += launch_reduce(B.pages, n, max_rh, max_rw, st);
    }
    CU(cudaEventRecord(L.ev[EV_CONV], st));
    if (!stream_in) {
        if (any_h) launches += launch_resample_h(B.pages, n, max_resh_rows, max_w, st);
        if (any_v) launches += launch_resample_v(B.pages, n, max_h, max_wc, st);
    }
    CU(cudaEventRecord(L.ev[EV_PIXEL], st));
    if (!stream_in) {
        launches += launch_png_filter(B.pages, n, max_h, max_wc, o.optimize, B.row_adler, B.row_busy, st);
        launches += launch_adler_combine(B.pages, n, B.row_adler, B.page_adler, st);
    } else {
        for (int i = 0; i < n; i++)     // stage-level: Adler-32 of a given stream (scratch: row_adler is unused here)
            launches += launch_adler_flat(A + plans[i].o_filt, (uint64_t)plans[i].filt_len,
                                          reinterpret_cast<uint32_t*>(A + o_tokens), B.page_adler + i, st);
    }
    CU(cudaEventRecord(L.ev[EV_FILTER], st));
    if (need_lz) { launches += launch_lz_order(B, st); launches += launch_lz(B, st); }
    CU(cudaEventRecord(L.ev[EV_LZ], st));
    if (need_huff) {
        launches += launch_huff_build(B, st);
        launches += launch_layout(B, st);
        launches += launch_payload_init(B, st);
        launches += launch_huff_emit(B, st);
    }
    CU(cudaEventRecord(L.ev[EV_HUFF], st));
    if (need_huff) launches += launch_png_finish(B, st);
    CU(cudaEventRecord(L.ev[EV_ASSEMBLE], st));
    if (need_huff) launches += launch_base64_pages(B, st);
    CU(cudaEventRecord(L.ev[EV_B64], st));
    CU(cudaGetLastError());
    // results block -> pinned host (after the descriptor mirror)
    uint8_t* HR = M + align_up(desc_bytes, 64);
    CU(cudaMemcpyAsync(HR, R, res_bytes, cudaMemcpyDeviceToHost, st));
    out.lane = &L; out.n = n; out.res_bytes = res_bytes; out.host_res = HR;
    out.launches = launches; out.in_bytes = in_bytes; out.filt_bytes = filt_bytes;
    out.d_png = B.png; out.d_b64 = B.b64; out.d_filt0 = A; out.d_tokens = B.tokens; out.d_sub_ntok = B.sub_ntok; out.d_sub_hist = B.sub_hist;
    out.nsub = nsub; out.nblocks = nblocks;
    return 0;
}

// Waits for a group's launch set and exposes its results block.
int finish_group(vcp_handle* h, GroupOut& out) {
    NvtxRange nvtx_("vcp wait launch set");
    Lane& L = *out.lane;
    const int n = out.n;
    CU(cudaStreamSynchronize(L.stream));
    uint8_t* HR = out.host_res;
    out.png_off = reinterpret_cast<const uint64_t*>(HR); out.png_len = out.png_off + n; out.b64_off = out.png_len + n; out.b64_len = out.b64_off + n;
    out.totals = out.b64_len + n;
    out.adler = reinterpret_cast<const uint32_t*>(out.totals + 2);
    const uint32_t err = *reinterpret_cast<const uint32_t*>(HR + out.res_bytes - 8);
    if (err) return fail(VCP_ESIZE, "internal output bound exceeded (flags %u)", err);
    h->stats.kernel_launches += out.launches;
    h->stats.in_bytes += out.in_bytes; h->stats.filtered_bytes += out.filt_bytes;
    float ms = 0;
    auto el = [&](int a, int b) { cudaEventElapsedTime(&ms, L.ev[a], L.ev[b]); return ms; };
    h->stats.ms_h2d += el(EV_START, EV_H2D); h->stats.ms_convert += el(EV_H2D, EV_CONV); h->stats.ms_resample += el(EV_CONV, EV_PIXEL); h->stats.ms_filter += el(EV_PIXEL, EV_FILTER);
    h->stats.ms_lz += el(EV_FILTER, EV_LZ); h->stats.ms_huff += el(EV_LZ, EV_HUFF); h->stats.ms_assemble += el(EV_HUFF, EV_ASSEMBLE);
    h->stats.ms_b64 += el(EV_ASSEMBLE, EV_B64);
    if (h->trace_t0) {                                    // VCP_TRACE: where this group's stages sit on the device clock of the call
        float t[EV_B64 + 1];
        for (int k = EV_START; k <= EV_B64; k++) { t[k] = 0; cudaEventElapsedTime(&t[k], h->trace_t0, L.ev[k]); }
        fprintf(stderr, "[vcp]   device: copy %.2f..%.2f pixel ..%.2f filter ..%.2f lz ..%.2f huff ..%.2f out ..%.2f\n",
                t[EV_START], t[EV_H2D], t[EV_PIXEL], t[EV_FILTER], t[EV_LZ], t[EV_HUFF], t[EV_B64]);
    }
    return 0;
}

int run_group(vcp_handle* h, std::vector<PagePlan>& plans, const vcp_opts& o, RunMode mode, GroupOut& out) {
    const int rc = issue_group(h, h->lane[0], plans, o, mode, out);
    return rc ? rc : finish_group(h, out);
}

void reset_stats(vcp_handle* h) { const uint64_t a = h->stats.arena_bytes; memset(&h->stats, 0, sizeof h->stats); h->stats.arena_bytes = a; }

}  // namespace

// ================================================================================================= public C ABI
extern "C" {

int vcp_version(void) { return VCP_VERSION; }
const char* vcp_last_error(void) { return g_err.c_str(); }

int vcp_init(int device, vcp_handle** out) {
    if (!out) return fail(VCP_EINVAL, "out is NULL");
    *out = nullptr;
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(VCP_EINVAL, "no CUDA device %d (%d visible)", device, ndev);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(VCP_ECUDA, "libvcprep is built for sm_100a (B200); device %d is sm_%d%d", device, prop.major, prop.minor);
    vcp_handle* h = new vcp_handle();
    h->device = device;
    for (Lane& L : h->lane) {
        cudaError_t e = cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete h; return fail(VCP_ECUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
        for (int i = 0; i < EV_COUNT; i++) cudaEventCreate(&L.ev[i]);
    }
    {
        const unsigned hc = std::thread::hardware_concurrency();
        int lw = 1;
        if (const char* g = getenv("LOCAL_WORLD_SIZE")) lw = std::max(1, atoi(g));
        h->copy_threads = std::max(2, std::min(16, (int)(hc ? hc : 8) / lw));   // B200 box, 16 vCPUs: 8 -> 16 threads = +7 % pages/s with PIL inputs
        if (const char* g = getenv("VCP_COPY_THREADS")) h->copy_threads = std::max(1, atoi(g));
    }
    if (const char* g = getenv("VCP_PIPE_BYTES")) { const long long v = atoll(g); if (v >= (1 << 20)) h->pipe_bytes = (size_t)v; }
    if (const char* g = getenv("VCP_GROUP_BYTES")) { const long long v = atoll(g); if (v >= (1 << 20)) h->group_bytes = (size_t)v; }
    *out = h;
    return 0;
}

void vcp_destroy(vcp_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    for (Lane& L : h->lane) {
        if (L.stream) cudaStreamSynchronize(L.stream);
        if (L.arena) cudaFree(L.arena);
        if (L.meta) cudaFreeHost(L.meta);
        if (L.stage) cudaFreeHost(L.stage);
        for (int i = 0; i < EV_COUNT; i++) if (L.ev[i]) cudaEventDestroy(L.ev[i]);
        if (L.stream) cudaStreamDestroy(L.stream);
    }
    if (h->trace_t0) cudaEventDestroy(h->trace_t0);
    delete h;
}

int vcp_check_page(const vcp_page_desc* page, const vcp_opts* opts) {
    if (!page || !opts) return fail(VCP_EINVAL, "bad arguments");
    if (opts->out_channels != 0 && opts->out_channels != 1 && opts->out_channels != 3) return fail(VCP_EINVAL, "out_channels must be 0, 1 or 3");
    PagePlan P;
    return plan_geometry(*page, *opts, P);
}

int vcp_output_bound(const vcp_page_desc* pages, int n, const vcp_opts* opts, uint64_t* png_bytes, uint64_t* b64_bytes) {
    if (n < 0 || (n > 0 && !pages) || !opts) return fail(VCP_EINVAL, "bad arguments");
    uint64_t p = 0, b = 0;
    for (int i = 0; i < n; i++) {
        PagePlan P;
        if (plan_geometry(pages[i], *opts, P) != 0) continue;      // a bad page produces no output
        p += P.png_bound; b += b64_bound_of(P.png_bound);
    }
    if (png_bytes) *png_bytes = p + 64;
    if (b64_bytes) *b64_bytes = opts->want_b64 ? b + 64 : 0;
    return 0;
}

}  // extern "C"

namespace {

struct Idat { const uint8_t* p; size_t n; };       // payload of one IDAT chunk (host memory)

// Accept or reject a page whose stream inflated without error, the way Image.open(png).load() would (returns the reason, or NULL).
// Pillow feeds zlib's inflate() one piece at a time — at most 64 KiB, never across IDAT chunks (PngImageFile.load_read with
// ImageFile.MAXBLOCK) — and stops as soon as the last row is out.  inflate() finishes everything that needs no output space in the
// call that produced that row: so a malformed block header there, or a wrong Adler-32 behind the final block, fails the image only
// when its bytes lie in the same piece as the end of the image; a trailer in a later piece is never looked at.  A stream whose final
// block ends before the image does, on a row boundary, is taken as it is when its Adler-32 is right (the missing rows stay black).
const char* decode_verdict(const DecPageD& R, const std::vector<Idat>& idats) {
    auto locate = [&](unsigned long long zoff, size_t* k, size_t* o) {
        size_t cum = 0;
        for (size_t q = 0; q < idats.size(); q++) { if (zoff < cum + idats[q].n) { *k = q; *o = (size_t)(zoff - cum); return true; } cum += idats[q].n; }
        return false;
    };
    auto piece_of = [&](unsigned long long zoff) -> long long { size_t k, o; return locate(zoff, &k, &o) ? (long long)((k << 24) | (o >> 16)) : -1; };
    auto trailer_at = [&](unsigned long long zoff, uint32_t* v) {
        uint32_t t = 0;
        for (int q = 0; q < 4; q++) { size_t k, o; if (!locate(zoff + q, &k, &o)) return false; t = (t << 8) | idats[k].p[o]; }
        *v = t; return true;
    };
    uint32_t want = 0;
    if (R.valid_len < R.filt_len) {
        // ZipDecode.c acts on the end of the stream only in a call that also completed a row: whole rows, and the final block with
        // its Adler-32 in the piece that completed the last of them
        const unsigned long long rowlen = 1ull + (unsigned long long)R.w * R.c;
        const unsigned long long a = (R.end_bit + 7) / 8;
        if (!R.end_bit || R.valid_len == 0 || R.valid_len % rowlen != 0 || !trailer_at(a, &want) ||
            piece_of(a + 3) != piece_of((R.end_bit - 1) >> 3)) return "image data truncated";
        return want == R.adler ? nullptr : "broken data stream (Adler-32 mismatch)";
    }
    if (!R.done_bit) return nullptr;                                  // the image ended inside a match: inflate() stopped right there
    const long long pc = piece_of((R.done_bit - 1) >> 3);
    if (R.post_err && piece_of(R.post_err_bit >> 3) == pc) return "broken data stream (malformed block behind the last row)";
    if (R.end_bit) {
        const unsigned long long a = (R.end_bit + 7) / 8;
        if (piece_of(a + 3) == pc && trailer_at(a, &want) && want != R.adler) return "broken data stream (Adler-32 mismatch)";
    }
    return nullptr;
}

// CRC-32 (IEEE 802.3, as zlib's crc32()) of a few host bytes: the chunk checks of the PNG container parse
uint32_t crc32_host(const uint8_t* p, size_t n) {
    static uint32_t tab[256];
    static std::once_flag once;
    std::call_once(once, [] {
        for (uint32_t i = 0; i < 256; i++) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1; tab[i] = c; }
    });
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; i++) c = tab[(c ^ p[i]) & 0xFFu] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}

int check_batch_args(vcp_handle* h, const vcp_page_desc* pages, int n, const vcp_opts* opts, void* out_png, void* out_b64, vcp_page_result* results) {
    if (!h || n < 0 || (n > 0 && (!pages || !results)) || !opts) return fail(VCP_EINVAL, "bad arguments");
    if (n > 0 && !out_png) return fail(VCP_EINVAL, "out_png is NULL");
    if (opts->want_b64 && n > 0 && !out_b64) return fail(VCP_EINVAL, "want_b64 set but out_b64 is NULL");
    if (opts->out_channels != 0 && opts->out_channels != 1 && opts->out_channels != 3) return fail(VCP_EINVAL, "out_channels must be 0, 1 or 3");
    return 0;
}

// The batch pipeline (caller holds h->mu).  `publish(first_page, last_page, lane)` is called once per group, in page order, right
// after the group's output bytes have been ENQUEUED to the caller's buffers on that lane's stream (results[] of its pages are final).
template <class Publish>
int run_batch(vcp_handle* h, const vcp_page_desc* pages, int n, const vcp_opts* opts,
              void* out_png, uint64_t png_cap, void* out_b64, uint64_t b64_cap, vcp_page_result* results, Publish publish) {
    CU(cudaSetDevice(h->device));
    reset_stats(h);
    // per-page validation: a bad page gets its own status and is left out of the launch set
    std::vector<PagePlan> all(n);
    for (int i = 0; i < n; i++) {
        memset(&results[i], 0, sizeof results[i]);
        const int rc = plan_geometry(pages[i], *opts, all[i]);
        all[i].status = rc;
        results[i].status = rc;
    }
    // ---- split into launch sets.  Device-resident sources: as large as group_bytes allows (fewer, fuller launches).
    //      Host sources: groups of ~pipe_bytes so that the H2D copy of group g+1 (other lane, other stream) overlaps the
    //      kernels of group g; results are collected in order.
    //      The first and last host groups are small (1/4, 1/2 of pipe_bytes): the first copy and the last launch set are the
    //      only parts of the pipeline nothing overlaps with.
    const size_t limit = opts->src_device ? h->group_bytes : std::min(h->group_bytes, h->pipe_bytes);
    std::vector<std::vector<int>> groups;
    {
        auto cost = [&](int i) { return opts->src_device ? (size_t)all[i].filt_len + (size_t)all[i].sw * all[i].sh * all[i].sc / 4
                                                         : (size_t)all[i].sw * all[i].sh * all[i].sc; };
        size_t rem = 0;
        for (int i = 0; i < n; i++) if (!all[i].status) rem += cost(i);
        size_t bytes = 0, lim = limit;
        for (int i = 0; i < n; i++) {
            if (all[i].status) continue;
            const size_t fb = cost(i);
            if (groups.empty() || (bytes + fb > lim && !groups.back().empty()) || (int)groups.back().size() >= kMaxGroupPages) {
                groups.emplace_back(); bytes = 0;
                lim = limit;
                if (!opts->src_device) {
                    const size_t g = groups.size() - 1;
                    if (g == 0) lim = limit / 4; else if (g == 1) lim = limit / 2;
                    if (rem <= limit - limit / 4) lim = std::min(lim, std::max(limit / 4, rem / 2));   // ramp down at the tail
                }
            }
            groups.back().push_back(i); bytes += fb; rem -= fb;
        }
    }
    const int G = (int)groups.size();
    std::vector<GroupOut> ctx(G);
    std::vector<std::vector<PagePlan>> gplans(G);
    uint64_t png_used = 0, b64_used = 0;
    const cudaMemcpyKind kind = opts->dst_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    auto issue = [&](int g) -> int {
        for (int i : groups[g]) gplans[g].push_back(all[i]);
        return issue_group(h, h->lane[g % kLanes], gplans[g], *opts, RUN_FULL, ctx[g]);
    };
    auto collect = [&](int g) -> int {          // wait for group g, place its bytes in the caller's buffers (async on its lane)
        GroupOut& go = ctx[g];
        int rc = finish_group(h, go);
        if (rc) return rc;
        Lane& L = *go.lane;
        const uint64_t gp = go.totals[0], gb = go.totals[1];
        if (png_used + gp > png_cap) return fail(VCP_ESIZE, "out_png too small: need %llu more bytes at offset %llu (cap %llu); size it with vcp_output_bound",
                                                 (unsigned long long)gp, (unsigned long long)png_used, (unsigned long long)png_cap);
        if (opts->want_b64 && b64_used + gb > b64_cap) return fail(VCP_ESIZE, "out_b64 too small (cap %llu)", (unsigned long long)b64_cap);
        CU(cudaEventRecord(L.ev[EV_B64], L.stream));
        if (gp) CU(cudaMemcpyAsync((uint8_t*)out_png + png_used, go.d_png, gp, kind, L.stream));
        if (opts->want_b64 && gb) CU(cudaMemcpyAsync((uint8_t*)out_b64 + b64_used, go.d_b64, gb, kind, L.stream));
        CU(cudaEventRecord(L.ev[EV_D2H], L.stream));
        for (size_t k = 0; k < groups[g].size(); k++) {
            vcp_page_result& r = results[groups[g][k]];
            const PagePlan& P = gplans[g][k];
            r.width = P.w; r.height = P.h; r.channels = P.c;
            r.png_off = png_used + go.png_off[k]; r.png_len = go.png_len[k];
            if (opts->want_b64) { r.b64_off = b64_used + go.b64_off[k]; r.b64_len = go.b64_len[k]; }
            r.adler32 = go.adler[k]; r.n_idat = (uint32_t)P.nblk;
            h->stats.png_bytes += r.png_len; h->stats.b64_bytes += r.b64_len;
        }
        png_used += gp; b64_used += gb;
        publish(groups[g].front(), groups[g].back(), &L);
        return 0;
    };
    int rc = 0;
    const bool trace = getenv("VCP_TRACE") != nullptr;
    // Host inputs, several groups: one call at a time feeds the host -> device link.  Two handles that copy at once (prepare_stream keeps
    // two batches in flight) each get half of it and BOTH finish late; taking turns, the second call's copies start when the first has
    // issued its last one, so the first call's drain (last group's kernels, payload copy, bytes) hides behind them (VCP_H2D_TURN=0: off).
    static std::mutex h2d_turn[kMaxDevices];
    static const bool use_turn = !(getenv("VCP_H2D_TURN") && atoi(getenv("VCP_H2D_TURN")) == 0);
    std::unique_lock<std::mutex> turn;
    if (use_turn && !opts->src_device && G > 1 && h->device >= 0 && h->device < kMaxDevices) turn = std::unique_lock<std::mutex>(h2d_turn[h->device]);
    if (trace) {
        if (!h->trace_t0) cudaEventCreate(&h->trace_t0);
        cudaEventRecord(h->trace_t0, h->lane[0].stream);
    } else if (h->trace_t0) { cudaEventDestroy(h->trace_t0); h->trace_t0 = nullptr; }
    const auto T0 = std::chrono::steady_clock::now();
    auto now_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - T0).count(); };
    for (int g = 0; g < G && !rc; g++) {
        // lane g % kLanes was last used by group g-kLanes: its payload copy must have left the arena before it is overwritten
        if (g >= kLanes) {
            Lane& L = h->lane[g % kLanes];
            const double ta = now_ms();
            CU(cudaStreamSynchronize(L.stream));
            if (trace) fprintf(stderr, "[vcp] g%d lane sync %.2f..%.2f\n", g, ta, now_ms());
            float ms = 0; cudaEventElapsedTime(&ms, L.ev[EV_B64], L.ev[EV_D2H]); h->stats.ms_d2h += ms;
        }
        const double tb = now_ms();
        rc = issue(g);
        const double tc = now_ms();
        if (!rc && g >= kLanes - 1) rc = collect(g - (kLanes - 1));
        if (trace) fprintf(stderr, "[vcp] g%d (%d pages) issue %.2f..%.2f collect(g%d) ..%.2f\n", g, (int)groups[g].size(), tb, tc, g - (kLanes - 1), now_ms());
    }
    if (turn.owns_lock()) turn.unlock();
    for (int g = std::max(0, G - (kLanes - 1)); g < G && !rc; g++) { rc = collect(g); if (trace) fprintf(stderr, "[vcp] tail collect(g%d) ..%.2f\n", g, now_ms()); }
    for (int l = 0; l < kLanes; l++) {             // drain both lanes even on error
        Lane& L = h->lane[l];
        cudaStreamSynchronize(L.stream);
        if (!rc && l < G) { float ms = 0; if (cudaEventElapsedTime(&ms, L.ev[EV_B64], L.ev[EV_D2H]) == cudaSuccess) h->stats.ms_d2h += ms; }
    }
    if (rc) return rc;
    h->stats.ms_total = h->stats.ms_h2d + h->stats.ms_convert + h->stats.ms_resample + h->stats.ms_filter + h->stats.ms_lz + h->stats.ms_huff +
                        h->stats.ms_assemble + h->stats.ms_b64 + h->stats.ms_d2h;
    return 0;
}

}  // namespace

extern "C" {

int vcp_prepare_batch(vcp_handle* h, const vcp_page_desc* pages, int n, const vcp_opts* opts,
                      void* out_png, uint64_t png_cap, void* out_b64, uint64_t b64_cap, vcp_page_result* results) {
    int rc = check_batch_args(h, pages, n, opts, out_png, out_b64, results);
    if (rc) return rc;
    LOCK_HANDLE(h);
    return run_batch(h, pages, n, opts, out_png, png_cap, out_b64, b64_cap, results, [](int, int, Lane*) {});
}

// ---- streaming variant: the pipeline runs on a worker thread of the handle; the caller picks up page ranges as their bytes land
int vcp_batch_begin(vcp_handle* h, const vcp_page_desc* pages, int n, const vcp_opts* opts,
                    void* out_png, uint64_t png_cap, void* out_b64, uint64_t b64_cap, vcp_page_result* results) {
    int rc = check_batch_args(h, pages, n, opts, out_png, out_b64, results);
    if (rc) return rc;
    BatchJob* J = new BatchJob();
    {
        std::lock_guard<std::mutex> lock(h->mu);
        if (h->job) { delete J; return fail(VCP_EINVAL, "a streaming batch is already in flight on this handle"); }
        h->job = J;                                 // from here to vcp_batch_end the handle belongs to the worker (see HandleGuard)
    }
    J->pages.assign(pages, pages + n); J->opts = *opts;
    try {
        J->th = std::thread([=]() {
            const int r = run_batch(h, J->pages.data(), n, &J->opts, out_png, png_cap, out_b64, b64_cap, results,
                                    [J](int first, int last, Lane* L) {
                                        cudaEvent_t ev = nullptr;
                                        cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
                                        cudaEventRecord(ev, L->stream);
                                        std::lock_guard<std::mutex> g(J->m);
                                        J->ready.push_back({first, last, ev});
                                        J->cv.notify_all();
                                    });
            std::lock_guard<std::mutex> g(J->m);
            J->rc = r; if (r) J->err = g_err;
            J->done = true;
            J->cv.notify_all();
        });
    } catch (...) {                                 // no worker: give the handle back
        std::lock_guard<std::mutex> lock(h->mu);
        h->job = nullptr;
        delete J;
        return fail(VCP_ENOMEM, "could not start the batch worker thread");
    }
    return 0;
}

int vcp_batch_next(vcp_handle* h, int* first_page, int* last_page) {
    if (!h || !h->job || !first_page || !last_page) return fail(VCP_EINVAL, "no batch in flight");
    BatchJob* J = h->job;
    BatchJob::Ready r;
    {
        std::unique_lock<std::mutex> g(J->m);
        J->cv.wait(g, [&] { return J->next < J->ready.size() || J->done; });
        if (J->next >= J->ready.size()) {                       // finished (or failed)
            if (J->rc) { g_err = J->err; return J->rc; }
            return 0;
        }
        r = J->ready[J->next++];
    }
    cudaSetDevice(h->device);
    const cudaError_t e = cudaEventSynchronize(r.ev);           // the group's bytes are now in the caller's buffers
    if (e != cudaSuccess) return fail(VCP_ECUDA, "cudaEventSynchronize: %s", cudaGetErrorString(e));
    *first_page = r.first; *last_page = r.last;
    return 1;
}

int vcp_batch_end(vcp_handle* h) {
    if (!h || !h->job) return fail(VCP_EINVAL, "no batch in flight");
    BatchJob* J = h->job;
    if (J->th.joinable()) J->th.join();
    for (auto& r : J->ready) if (r.ev) cudaEventDestroy(r.ev);
    const int rc = J->rc;
    if (rc) g_err = J->err;
    delete J;
    {
        std::lock_guard<std::mutex> lock(h->mu);
        h->job = nullptr;
    }
    return rc;
}

// ------------------------------------------------------------------------------------------ PNG decode
int vcp_png_decode_batch(vcp_handle* h, const void* const* pngs, const uint64_t* png_lens, int n,
                         void* out_pixels, uint64_t out_cap, int dst_device, vcp_decode_result* results) {
    if (!h || n < 0 || (n > 0 && (!pngs || !png_lens || !results || !out_pixels))) return fail(VCP_EINVAL, "bad arguments");
    NvtxRange nvtx_("vcp_png_decode_batch");
    LOCK_HANDLE(h);
    CU(cudaSetDevice(h->device));
    std::vector<std::vector<Idat>> idats(n);
    std::vector<DecPageD> dp(n);
    auto be32 = [](const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; };
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    // ---- container parse on the host, as PngImagePlugin does it (PIL/PngImagePlugin.py: PngImageFile._open, load_read, load_end):
    //      every chunk in front of the first IDAT must be whole and carry a correct CRC-32 (ChunkStream.crc); the image data are the
    //      CONSECUTIVE IDAT chunks from there (their CRCs are skipped, not checked); an IDAT that runs past the end of the file gives
    //      what is there; whatever follows the data (IEND, or nothing at all) is not needed.
    for (int i = 0; i < n; i++) {
        memset(&results[i], 0, sizeof results[i]);
        DecPageD& D = dp[i]; memset(&D, 0, sizeof D);
        const uint8_t* p = (const uint8_t*)pngs[i]; const size_t len = (size_t)png_lens[i];
        int st = VCP_OK;
        if (!p || len < 8 || memcmp(p, sig, 8) != 0) st = fail(VCP_EINVAL, "PNG %d: not a PNG", i);
        size_t off = 8; bool seen_hdr = false, in_data = false;
        auto is_cid = [](const uint8_t* t) { for (int k = 0; k < 4; k++) if (!(isalnum(t[k]) || t[k] == '_')) return false; return true; };
        while (st == VCP_OK) {
            if (off + 8 > len) {                                       // no further chunk header
                if (!in_data) st = fail(VCP_EINVAL, "PNG %d: truncated before the image data", i);
                break;
            }
            const size_t cl = be32(p + off);
            const uint8_t* tag = p + off + 4; const uint8_t* data = p + off + 8;
            if (!is_cid(tag)) {
                // in front of the data: "broken PNG file (chunk ...)".  Behind an IDAT Pillow meets the bad header only if its decoder
                // still wants bytes — then the image is incomplete and fails here as well; if it is complete nothing reads the header.
                if (!in_data) st = fail(VCP_EINVAL, "PNG %d: broken chunk header", i);
                break;
            }
            if (!memcmp(tag, "IDAT", 4)) {
                if (!seen_hdr) { st = fail(VCP_EINVAL, "PNG %d: IDAT before IHDR", i); break; }
                in_data = true;
                const size_t have = std::min(cl, len - (off + 8));
                if (have) idats[i].push_back({data, have});
                if (have < cl) break;                                  // the file ends inside this chunk
                off += 12 + cl;
                continue;
            }
            if (in_data) break;                                        // the first other chunk ends the image data
            if (off + 12 + cl > len) { st = fail(VCP_EINVAL, "PNG %d: truncated chunk", i); break; }
            if (crc32_host(tag, 4 + cl) != be32(data + cl)) { st = fail(VCP_EINVAL, "PNG %d: bad checksum in chunk %.4s", i, (const char*)tag); break; }
            if (!memcmp(tag, "IHDR", 4)) {
                if (cl != 13) { st = fail(VCP_EINVAL, "PNG %d: bad IHDR", i); break; }
                D.w = (int32_t)be32(data); D.h = (int32_t)be32(data + 4);
                const int depth = data[8], ct = data[9], il = data[12];
                D.c = ct == 0 ? 1 : ct == 4 ? 2 : ct == 2 ? 3 : ct == 6 ? 4 : 0;
                if (depth != 8 || !D.c || il != 0 || data[10] != 0 || data[11] != 0 || D.w <= 0 || D.h <= 0 ||
                    (int64_t)D.w * D.c > (1 << 24) || D.h > (1 << 24))
                    { st = fail(VCP_EINVAL, "PNG %d: only 8-bit gray/gray+alpha/RGB/RGBA non-interlaced images are on the path", i); break; }
                seen_hdr = true;
            } else if (!memcmp(tag, "IEND", 4)) { st = fail(VCP_EINVAL, "PNG %d: no image data", i); break; }
            off += 12 + cl;
        }
        if (st == VCP_OK && (!seen_hdr || idats[i].empty())) st = fail(VCP_EINVAL, "PNG %d: no image data", i);
        if (st == VCP_OK) {
            D.filt_len = (unsigned long long)D.h * (1ull + (unsigned long long)D.w * D.c);
            if (D.filt_len >= (1ull << 32)) st = fail(VCP_EINVAL, "PNG %d: too large", i);
        }
        results[i].status = st;
        D.status = st;
    }
    // ---- the batch runs as up to two groups of consecutive pages on two lanes (streams): the latency-bound kernels of one group
    //      (probe, window chain, un-filter pipeline) overlap the throughput-bound ones of the other (exec, resolve)
    constexpr unsigned long long kResolveChunk = 32768, kCkpt = 65536, kScanBits = 8192;      // keep in step with png_decode.cu
    constexpr int kMaxCand = 1024;                                                              // scanned block headers per page
    constexpr unsigned long long kSpecBitsMax = VCP_SPEC_BITS, kSpecBitsMin = 8192, kSpecUnits = 3000;   // speculative start points (png_decode.cu: k_infl_spec)
    struct Group { int i0 = 0, i1 = 0; uint64_t pix_base = 0, pix_bytes = 0; DecPageD* hd = nullptr; Lane* L = nullptr; const DecSegD* d_segs = nullptr; int seg_total = 0; };
    auto enqueue = [&](Group& G) -> int {
        Lane& L = *G.L;
        const int i0 = G.i0, m = G.i1 - G.i0;
        Bump bump; size_t zoff_total = 0;
        std::vector<size_t> o_f(m, kNone), o_p(m, kNone), o_s(m, kNone), s_z(m, 0);
        std::vector<unsigned long long> cand;            // candidate start bits: per page seg_cap entries, the IDAT starts filled in
        std::vector<uint32_t> chunk_page, chunk_pos, scan_page, scan_bit;
        uint64_t pix_total = 0;
        int nbands = 0;
        size_t nslots = 0, iv_total = 0, seg_total = 0, surv_total = 0, spec_total = 0;
        // One speculative parse start per spec_bits of a page's stream: 4 KiB when the batch is large (its first parse then has thousands of
        // units, more than the GPU holds at once), down to 1 KiB when a few pages with short streams would otherwise leave it a few
        // hundred serial units on 148 SMs.  Literal-heavy pages (long streams) keep the coarse spacing: every start costs a 6 Kbit walk.
        int m_ok = 0;
        for (int j = 0; j < m; j++) m_ok += !dp[i0 + j].status;
        const unsigned long long units_per_page = std::max<unsigned long long>(64, kSpecUnits / (unsigned long long)std::max(1, m_ok));
        for (int j = 0; j < m; j++) {
            DecPageD& D = dp[i0 + j];
            D.band0 = nbands; D.iv0 = (int32_t)iv_total; D.seg0 = D.cand0 = (int32_t)seg_total; D.surv0 = (int32_t)surv_total; D.slot0 = (int32_t)nslots;
            if (D.status) continue;
            size_t zl = 0; for (auto& c : idats[i0 + j]) zl += c.n;
            if (zl >= (1ull << 28) || idats[i0 + j].size() > (1u << 20)) { results[i0 + j].status = D.status = fail(VCP_EINVAL, "PNG %d: too large", i0 + j); continue; }
            D.zlen = zl;
            s_z[j] = zoff_total; zoff_total += align_up(zl + 64, 256);
            o_f[j] = bump.take((size_t)D.filt_len + 16 + 512);          // k_unfilter's TMA rows read up to 62 pixels + a granule past the last row
            o_s[j] = bump.take(((size_t)D.filt_len + 16) * 2);
            nbands += (D.h + 31) / 32;
            // parse units: every IDAT start, plus up to kMaxCand block headers the scan finds on the device.  A parse can produce at
            // most the whole page: page_iv checkpoint slots each.
            const unsigned long long page_iv = D.filt_len / kCkpt + 3;
            D.page_iv = (int32_t)page_iv;
            int n_idat = 0; size_t o = 0, q = 0;
            const size_t keep_every = (idats[i0 + j].size() + 2047) / 2048;       // parse units are optional: a PNG cut into very many IDATs offers a subset
            for (auto& c : idats[i0 + j]) {
                const unsigned long long sb = o == 0 ? 16ull : 8ull * o;         // the deflate data starts behind the 2-byte zlib header
                if ((n_idat == 0 || sb > cand.back()) && q % keep_every == 0) { cand.push_back(sb); n_idat++; }
                o += c.n; q++;
            }
            D.n_idat = n_idat; D.ncand = (uint32_t)n_idat; D.nsurv = 0;
            unsigned long long spec_bits = kSpecBitsMax;
            while (spec_bits > kSpecBitsMin && zl * 8ull / spec_bits < units_per_page) spec_bits >>= 1;
            D.spec_bits = (uint32_t)spec_bits;
            D.spec0 = (int32_t)spec_total; D.nspec = (int32_t)(zl * 8ull / spec_bits);
            spec_total += (size_t)D.nspec;
            D.seg_cap = n_idat + kMaxCand + D.nspec;
            cand.resize(seg_total + (size_t)D.seg_cap, 0ull);
            seg_total += (size_t)D.seg_cap;
            nslots += (size_t)D.seg_cap * page_iv;
            D.surv_cap = (int32_t)(zl * 8 / 32 + 256);
            surv_total += (size_t)D.surv_cap;
            // An encoder that ends every IDAT on a block boundary by an empty stored block (this library: 00 00 FF FF closes every IDAT but
            // the last) has its block starts at the IDAT starts already; the chain check in k_infl_plan still decides, the scan is skipped.
            bool synced = idats[i0 + j].size() >= 2;
            for (size_t q = 0; synced && q + 1 < idats[i0 + j].size(); q++) {
                const Idat& c = idats[i0 + j][q];
                synced = c.n >= 4 && c.p[c.n - 4] == 0 && c.p[c.n - 3] == 0 && c.p[c.n - 2] == 0xFF && c.p[c.n - 1] == 0xFF;
            }
            if (!synced || getenv("VCP_DECODE_SCAN_ALL"))
                for (unsigned long long bb = 0; bb < zl * 8ull; bb += kScanBits) { scan_page.push_back((uint32_t)j); scan_bit.push_back((uint32_t)bb); }
            D.iv_cap = (int32_t)(page_iv + (unsigned long long)D.seg_cap);
            iv_total += (size_t)D.iv_cap;
            D.chunk0 = (int32_t)chunk_page.size();
            for (unsigned long long p = 0; p < D.filt_len; p += kResolveChunk) { chunk_page.push_back((uint32_t)j); chunk_pos.push_back((uint32_t)p); }
        }
        if (nslots >= (1ull << 31) || iv_total >= (1ull << 31) || surv_total >= (1ull << 31)) return fail(VCP_ESIZE, "PNG batch too large");
        const size_t o_zreg = bump.take(zoff_total + 256);
        const size_t o_preg = bump.take(0);
        for (int j = 0; j < m; j++) {
            DecPageD& D = dp[i0 + j];
            if (D.status) continue;
            const size_t pl = (size_t)D.w * D.h * D.c;
            o_p[j] = bump.take(pl);
            vcp_decode_result& R = results[i0 + j];
            R.pix_off = G.pix_base + (o_p[j] - o_preg); R.pix_len = pl;
            R.width = D.w; R.height = D.h; R.channels = D.c;
            pix_total = (o_p[j] - o_preg) + pl;
        }
        G.pix_bytes = align_up(pix_total, 256);
        // un-filter work list: bands g .. g + VCP_UF_GROUP - 1 of every page that has them, for g = 0, VCP_UF_GROUP, ...
        std::vector<uint32_t> band_page, band_idx;
        band_page.reserve(nbands); band_idx.reserve(nbands);
        {
            int maxb = 0;
            for (int j = 0; j < m; j++) if (!dp[i0 + j].status) maxb = std::max(maxb, (dp[i0 + j].h + 31) / 32);
            for (int bnd = 0; bnd < maxb; bnd += VCP_UF_GROUP)
                for (int j = 0; j < m; j++) if (!dp[i0 + j].status && bnd < (dp[i0 + j].h + 31) / 32) { band_page.push_back((uint32_t)j); band_idx.push_back((uint32_t)bnd); }
        }
        const size_t desc_bytes = align_up((size_t)m * sizeof(DecPageD), 256), seg_bytes = align_up(cand.size() * sizeof(unsigned long long), 256),
                     chunk_bytes = align_up(chunk_page.size() * sizeof(uint32_t), 256), band_bytes = align_up(band_page.size() * sizeof(uint32_t), 256),
                     scan_bytes = align_up(scan_page.size() * sizeof(uint32_t), 256);
        const size_t meta_bytes = desc_bytes + seg_bytes + 2 * chunk_bytes + 2 * band_bytes + 2 * scan_bytes;
        const size_t o_chdr = bump.take(seg_bytes + 16);                  // candidate header bits (zeroed on the device)
        const size_t o_desc = bump.take(meta_bytes + 16);
        const size_t flag_bytes = align_up(((size_t)nbands + 64) * sizeof(uint32_t), 256);
        const size_t o_flag = bump.take(flag_bytes);
        const size_t o_slots = bump.take((nslots + 1) * sizeof(DecIvD)), o_ivs = bump.take((iv_total + 1) * sizeof(DecIvD));
        const size_t o_segs = bump.take((seg_total + 1) * sizeof(DecSegD)), o_surv = bump.take((surv_total + 1) * sizeof(uint32_t));
        const size_t o_cadl = bump.take((chunk_page.size() + 1) * 2 * sizeof(uint32_t));
        if (G.pix_base + pix_total > out_cap) return fail(VCP_ESIZE, "out_pixels too small: need %llu bytes", (unsigned long long)(G.pix_base + pix_total));
        // device output: the un-filter writes the caller's buffer directly (it reads whole words, so the buffer must hold the last page's padding)
        const bool direct = dst_device && G.pix_base + align_up(pix_total, 4) <= out_cap;
        int rc = ensure_arena(L, bump.off + 256); if (rc) return rc;
        rc = ensure_stage(L, zoff_total + meta_bytes + 512); if (rc) return rc;
        uint8_t* A = L.arena;
        std::vector<CopyJob> jobs;
        for (int j = 0; j < m; j++) {
            DecPageD& D = dp[i0 + j];
            if (D.status) continue;
            size_t o = s_z[j];
            for (auto& c : idats[i0 + j]) { jobs.push_back({L.stage + o, c.p, c.n, false}); o += c.n; }
            D.z = A + o_zreg + s_z[j]; D.filt = A + o_f[j];
            D.pix = direct ? (uint8_t*)out_pixels + G.pix_base + (o_p[j] - o_preg) : A + o_p[j];
            D.sym = reinterpret_cast<uint16_t*>(A + o_s[j]);
        }
        parallel_copy(jobs, h->copy_threads);
        uint8_t* hm = L.stage + align_up(zoff_total, 256);
        G.hd = reinterpret_cast<DecPageD*>(hm);
        memcpy(G.hd, dp.data() + i0, (size_t)m * sizeof(DecPageD));
        if (!cand.empty()) memcpy(hm + desc_bytes, cand.data(), cand.size() * sizeof(unsigned long long));
        if (!chunk_page.empty()) {
            memcpy(hm + desc_bytes + seg_bytes, chunk_page.data(), chunk_page.size() * sizeof(uint32_t));
            memcpy(hm + desc_bytes + seg_bytes + chunk_bytes, chunk_pos.data(), chunk_pos.size() * sizeof(uint32_t));
        }
        if (!band_page.empty()) {
            memcpy(hm + desc_bytes + seg_bytes + 2 * chunk_bytes, band_page.data(), band_page.size() * sizeof(uint32_t));
            memcpy(hm + desc_bytes + seg_bytes + 2 * chunk_bytes + band_bytes, band_idx.data(), band_idx.size() * sizeof(uint32_t));
        }
        const size_t scan_off = desc_bytes + seg_bytes + 2 * chunk_bytes + 2 * band_bytes;
        if (!scan_page.empty()) {
            memcpy(hm + scan_off, scan_page.data(), scan_page.size() * sizeof(uint32_t));
            memcpy(hm + scan_off + scan_bytes, scan_bit.data(), scan_bit.size() * sizeof(uint32_t));
        }
        cudaStream_t st = L.stream;
        if (zoff_total) CU(cudaMemcpyAsync(A + o_zreg, L.stage, zoff_total, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(A + o_desc, hm, meta_bytes, cudaMemcpyHostToDevice, st));
        CU(cudaMemsetAsync(A + o_flag, 0, flag_bytes, st));
        CU(cudaMemsetAsync(A + o_chdr, 0, seg_bytes, st));
        DecBatchD B; memset(&B, 0, sizeof B);
        B.pages = reinterpret_cast<DecPageD*>(A + o_desc); B.npages = m;
        B.segs = reinterpret_cast<DecSegD*>(A + o_segs); B.seg_total = (int32_t)seg_total;
        G.d_segs = B.segs; G.seg_total = (int)seg_total;
        B.cand_bits = reinterpret_cast<unsigned long long*>(A + o_desc + desc_bytes);
        B.cand_hdr = reinterpret_cast<unsigned long long*>(A + o_chdr); B.spec_total = (int32_t)spec_total;
        B.surv = reinterpret_cast<uint32_t*>(A + o_surv); B.surv_total = (int32_t)surv_total;
        B.scan_page = reinterpret_cast<const uint32_t*>(A + o_desc + scan_off);
        B.scan_bit = reinterpret_cast<const uint32_t*>(A + o_desc + scan_off + scan_bytes);
        B.nscan = (int32_t)scan_page.size();
        if (getenv("VCP_DECODE_NO_SCAN")) B.no_scan = 1;
        B.slots = reinterpret_cast<DecIvD*>(A + o_slots);
        B.ivs = reinterpret_cast<DecIvD*>(A + o_ivs); B.iv_total = (int32_t)iv_total;
        B.chunk_page = reinterpret_cast<const uint32_t*>(A + o_desc + desc_bytes + seg_bytes);
        B.chunk_pos = reinterpret_cast<const uint32_t*>(A + o_desc + desc_bytes + seg_bytes + chunk_bytes);
        B.nchunks = (int32_t)chunk_page.size();
        B.chunk_adler = reinterpret_cast<uint32_t*>(A + o_cadl);
        B.counters = reinterpret_cast<uint32_t*>(A + o_flag);
        B.band_flag = B.counters + 64; B.nbands = nbands; B.ngroups = (int32_t)band_page.size();
        B.band_page = reinterpret_cast<const uint32_t*>(A + o_desc + desc_bytes + seg_bytes + 2 * chunk_bytes);
        B.band_idx = reinterpret_cast<const uint32_t*>(A + o_desc + desc_bytes + seg_bytes + 2 * chunk_bytes + band_bytes);
        if (const char* g = getenv("VCP_DBG_UF_NOWAIT")) B.dbg_nowait = atoi(g);
        launch_inflate(B, st);
        launch_unfilter(B, st);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(G.hd, B.pages, (size_t)m * sizeof(DecPageD), cudaMemcpyDeviceToHost, st));
        if (pix_total && !direct) CU(cudaMemcpyAsync((uint8_t*)out_pixels + G.pix_base, A + o_preg, pix_total, dst_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
        return 0;
    };
    // split where half of the filtered bytes have gone by.  Measured on B200: +5 % at 256 letter pages, a loss at 64 (two half-sized
    // latency-bound pipelines instead of one), so small batches stay whole
    unsigned long long total_f = 0; int n_ok = 0;
    for (int i = 0; i < n; i++) if (!dp[i].status) { total_f += dp[i].filt_len; n_ok++; }
    int split = n;
    if (n_ok >= 128 && !getenv("VCP_DECODE_ONE_GROUP")) {
        unsigned long long acc = 0;
        for (int i = 0; i < n; i++) { if (!dp[i].status) acc += dp[i].filt_len; if (acc * 2 >= total_f) { split = i + 1; break; } }
        if (split >= n) split = n;
    }
    Group G[2];
    G[0].i0 = 0; G[0].i1 = split; G[0].L = &h->lane[0];
    G[1].i0 = split; G[1].i1 = n; G[1].L = &h->lane[1];
    int rc = 0;
    for (int g = 0; g < 2 && rc == 0; g++) {
        if (G[g].i1 <= G[g].i0) continue;
        if (g == 1) G[1].pix_base = G[0].pix_base + G[0].pix_bytes;
        rc = enqueue(G[g]);
    }
    for (int g = 0; g < 2; g++) {
        if (G[g].i1 <= G[g].i0) continue;
        cudaError_t e = cudaStreamSynchronize(G[g].L->stream);
        if (e != cudaSuccess && rc == 0) rc = fail(VCP_ECUDA, "CUDA: %s", cudaGetErrorString(e));
    }
    if (rc) return rc;
    if (getenv("VCP_DECODE_DEBUG"))
        for (int g = 0; g < 2; g++) if (G[g].seg_total > 0) {
            std::vector<DecSegD> hs((size_t)G[g].seg_total);
            cudaMemcpy(hs.data(), G[g].d_segs, hs.size() * sizeof(DecSegD), cudaMemcpyDeviceToHost);
            const DecPageD& R = G[g].hd[0];
            fprintf(stderr, "[vcp] parse units of page %d (start bit, header bit, bytes out, ok, kilo-cycles):", G[g].i0);
            for (int k = 0; k < R.nseg && k < 90; k++) { const DecSegD& S = hs[(size_t)R.seg0 + k];
                fprintf(stderr, " (%llu,%llu,%u,%d,%u)", (unsigned long long)S.start_bit, (unsigned long long)S.hdr_bit, S.olen, S.ok, S.pad); }
            fprintf(stderr, "\n");
        }
    for (int g = 0; g < 2; g++)
        for (int i = G[g].i0; i < G[g].i1; i++) {
            if (!G[g].hd) continue;
            const DecPageD& R = G[g].hd[i - G[g].i0];
            if (getenv("VCP_DECODE_DEBUG"))
                fprintf(stderr, "[vcp] decode page %d: zlen %llu n_idat %d nspec %d candidates %u parse units %d intervals %d status %d\n", i,
                        (unsigned long long)R.zlen, R.n_idat, R.nspec, R.ncand, R.nseg, R.niv, R.status);
            if (results[i].status) continue;
            if (R.status) { results[i].status = VCP_EINVAL; fail(VCP_EINVAL, "PNG %d: corrupt zlib stream or filter byte (code %d)", i, R.status); continue; }
            const char* why = decode_verdict(R, idats[i]);
            if (why) { results[i].status = VCP_EINVAL; fail(VCP_EINVAL, "PNG %d: %s", i, why); }
        }
    return 0;
}

int vcp_get_stats(vcp_handle* h, vcp_stats* out) {
    if (!h || !out) return fail(VCP_EINVAL, "bad arguments");
    LOCK_HANDLE(h);
    *out = h->stats;
    return 0;
}

int vcp_host_scatter(const void* src_base, const uint64_t* offs, const uint64_t* lens, void* const* dsts, int n, int threads) {
    if (n < 0 || (n > 0 && (!src_base || !offs || !lens || !dsts))) return fail(VCP_EINVAL, "bad arguments");
    const uint8_t* base = (const uint8_t*)src_base;
    uint64_t total = 0;
    for (int i = 0; i < n; i++) total += lens[i];
    const int T = std::max(1, std::min(threads, 16));
    if (T == 1 || total < (1u << 20)) {
        for (int i = 0; i < n; i++) if (lens[i]) memcpy(dsts[i], base + offs[i], (size_t)lens[i]);
        return 0;
    }
    // split the total byte count evenly: part t copies the byte interval [t*total/T, (t+1)*total/T) of the concatenation
    HostPool::get().run(T, T, [=](int t) {
        const uint64_t lo = total * t / T, hi = total * (t + 1) / T;
        uint64_t pos = 0;
        for (int i = 0; i < n && pos < hi; i++) {
            const uint64_t a = std::max(lo, pos), b = std::min(hi, pos + lens[i]);
            if (a < b) memcpy((uint8_t*)dsts[i] + (a - pos), base + offs[i] + (a - pos), (size_t)(b - a));
            pos += lens[i];
        }
    });
    return 0;
}

// ------------------------------------------------------------------------------------------ stage-level entry points
namespace {
// upload one PageD built on the host and return its device address (uses the head of the arena's descriptor area)
int stage_page(Lane& L, const PageD& D, size_t extra_bytes, uint8_t** extra, const PageD** d_page) {
    const size_t need = 1024 + align_up(extra_bytes, 256) + 512;
    int rc = ensure_arena(L, need); if (rc) return rc;
    rc = ensure_meta(L, 4096); if (rc) return rc;
    memcpy(L.meta, &D, sizeof D);
    CU(cudaMemcpyAsync(L.arena, L.meta, sizeof D, cudaMemcpyHostToDevice, L.stream));
    *d_page = reinterpret_cast<const PageD*>(L.arena);
    if (extra) *extra = L.arena + 1024;
    return 0;
}
}  // namespace

int vcp_convert(vcp_handle* h, const void* d_src, int width, int height, int src_channels, int64_t row_stride,
                void* d_dst, int dst_channels) {
    if (!h || !d_src || !d_dst || width <= 0 || height <= 0) return fail(VCP_EINVAL, "bad arguments");
    if (src_channels < 1 || src_channels > 4 || (dst_channels != 1 && dst_channels != 3)) return fail(VCP_EINVAL, "unsupported conversion %d -> %d channels", src_channels, dst_channels);
    LOCK_HANDLE(h);
    CU(cudaSetDevice(h->device));
    Lane& L = h->lane[0]; (void)L;
    PageD D = {};
    D.src = (const uint8_t*)d_src; D.src_stride = row_stride ? row_stride : (int64_t)width * src_channels;
    D.sw = width; D.sh = height; D.sc = src_channels; D.c = dst_channels; D.pc = dst_channels; D.conv = (uint8_t*)d_dst;
    const PageD* dp; int rc = stage_page(L, D, 0, nullptr, &dp); if (rc) return rc;
    launch_convert(dp, 1, height, width, L.stream);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(L.stream));
    return 0;
}

int vcp_resample_coeffs(int in_size, int out_size, int filter, float box0, float box1, int32_t* bounds, int32_t* kk, int* ksize) {
    if (resample_coeffs_host(in_size, out_size, filter, box0, box1, bounds, kk, ksize) != 0) return fail(VCP_EINVAL, "bad resample parameters");
    return 0;
}

int vcp_reduce(vcp_handle* h, const void* d_src, int width, int height, int channels, void* d_dst, int fx, int fy) {
    if (!h || !d_src || !d_dst || width <= 0 || height <= 0 || channels < 1 || channels > 4 || fx < 1 || fy < 1) return fail(VCP_EINVAL, "bad arguments");
    LOCK_HANDLE(h);
    CU(cudaSetDevice(h->device));
    Lane& L = h->lane[0]; (void)L;
    PageD D = {};
    D.sw = width; D.sh = height; D.sc = channels; D.c = channels; D.pc = channels; D.fx = fx; D.fy = fy;
    D.rw = (width + fx - 1) / fx; D.rh = (height + fy - 1) / fy;
    D.rdin = (const uint8_t*)d_src; D.rdin_stride = (int64_t)width * channels; D.red = (uint8_t*)d_dst;
    const PageD* dp; int rc = stage_page(L, D, 0, nullptr, &dp); if (rc) return rc;
    launch_reduce(dp, 1, D.rh, D.rw, L.stream);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(L.stream));
    return 0;
}

int vcp_resample(vcp_handle* h, const void* d_src, int width, int height, int channels,
                 void* d_dst, int out_width, int out_height, int filter) {
    if (!h || !d_src || !d_dst || width <= 0 || height <= 0 || out_width <= 0 || out_height <= 0) return fail(VCP_EINVAL, "bad arguments");
    if (channels != 1 && channels != 3) return fail(VCP_EINVAL, "resample supports L and RGB");
    if (filter < VCP_LANCZOS || filter > VCP_HAMMING) return fail(VCP_EINVAL, "unsupported resample filter %d", filter);
    LOCK_HANDLE(h);
    CU(cudaSetDevice(h->device));
    Lane& L = h->lane[0]; (void)L;
    const bool need_h = out_width != width, need_v = out_height != height;
    if (!need_h && !need_v) {
        CU(cudaMemcpyAsync(d_dst, d_src, (size_t)width * height * channels, cudaMemcpyDeviceToDevice, L.stream));
        CU(cudaStreamSynchronize(L.stream));
        return 0;
    }
    const CoeffRef ch = need_h ? get_coeffs(h, width, out_width, filter, 0.f, (float)width) : nullptr;
    const CoeffRef cv = need_v ? get_coeffs(h, height, out_height, filter, 0.f, (float)height) : nullptr;
    if ((need_h && !ch) || (need_v && !cv)) return fail(VCP_EINVAL, "bad resample parameters");
    std::vector<int32_t> blob;
    size_t ihb = 0, ihk = 0, ivb = 0, ivk = 0;
    if (ch) { ihb = blob.size(); blob.insert(blob.end(), ch->bounds.begin(), ch->bounds.end()); ihk = blob.size(); blob.insert(blob.end(), ch->kt.begin(), ch->kt.end()); }
    if (cv) { ivb = blob.size(); blob.insert(blob.end(), cv->bounds.begin(), cv->bounds.end()); ivk = blob.size(); blob.insert(blob.end(), cv->kt.begin(), cv->kt.end()); }
    const size_t tmp_bytes = (need_h && need_v) ? (size_t)out_width * height * channels : 0;
    const size_t extra = align_up(blob.size() * 4, 256) + tmp_bytes + 256;
    int rc = ensure_arena(L, 1024 + extra + 512); if (rc) return rc;
    int32_t* d_blob = reinterpret_cast<int32_t*>(L.arena + 1024);
    uint8_t* d_tmp = L.arena + 1024 + align_up(blob.size() * 4, 256);
    CU(cudaMemcpyAsync(d_blob, blob.data(), blob.size() * 4, cudaMemcpyHostToDevice, L.stream));
    CU(cudaStreamSynchronize(L.stream));      // blob is pageable host memory
    PageD D = {};
    D.c = channels; D.pc = channels; D.rw = width; D.rh = height; D.w = out_width; D.h = out_height;
    D.hin = (const uint8_t*)d_src; D.hin_stride = (int64_t)width * channels;
    const uint8_t* cur = D.hin; int64_t cur_stride = D.hin_stride;
    if (need_h) {
        D.tmp = need_v ? d_tmp : (uint8_t*)d_dst; D.hb = d_blob + ihb; D.hk = d_blob + ihk; D.hks = ch->ksize;
        D.tmp_stride = (int64_t)out_width * channels;
        cur = D.tmp; cur_stride = D.tmp_stride;
    }
    D.vin = cur; D.vin_stride = cur_stride;
    if (need_v) { D.vout = (uint8_t*)d_dst; D.vout_stride = (int64_t)out_width * channels; D.vb = d_blob + ivb; D.vk = d_blob + ivk; D.vks = cv->ksize; }
    const PageD* dp; rc = stage_page(L, D, 0, nullptr, &dp); if (rc) return rc;
    if (need_h) launch_resample_h(dp, 1, height, out_width, L.stream);
    if (need_v) launch_resample_v(dp, 1, out_height, out_width * channels, L.stream);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(L.stream));
    return 0;
}

int vcp_png_filter(vcp_handle* h, const void* d_pix, int width, int height, int channels, int optimize,
                   void* d_dst, uint32_t* adler32_out) {
    if (!h || !d_pix || !d_dst || width <= 0 || height <= 0 || channels < 1 || channels > 4) return fail(VCP_EINVAL, "bad arguments");
    LOCK_HANDLE(h);
    CU(cudaSetDevice(h->device));
    Lane& L = h->lane[0]; (void)L;
    vcp_page_desc d = {}; d.src = d_pix; d.width = width; d.height = height; d.channels = channels;
    vcp_opts o = {}; o.out_channels = 0; o.optimize = optimize; o.src_device = 1; o.compress_level = 6;
    std::vector<PagePlan> g(1);
    int rc = plan_geometry(d, o, g[0]); if (rc) return rc;
    GroupOut go;
    rc = run_group(h, g, o, RUN_FILTER_ONLY, go); if (rc) return rc;
    CU(cudaMemcpyAsync(d_dst, L.arena + go.filt_off[0], (size_t)g[0].filt_len, cudaMemcpyDeviceToDevice, L.stream));
    CU(cudaStreamSynchronize(L.stream));
    if (adler32_out) *adler32_out = go.adler[0];
    return 0;
}

static int stream_plan(const void* d_stream, uint64_t len, int bpp, PagePlan& P) {
    if (!d_stream || len == 0 || len >= ((uint64_t)1 << 31) - (1 << 20)) return fail(VCP_EINVAL, "bad stream length %llu", (unsigned long long)len);
    if (bpp < 1 || bpp > 4) return fail(VCP_EINVAL, "bpp must be 1..4");
    P = PagePlan();
    P.src = (const uint8_t*)d_stream; P.c = bpp; P.pc = bpp; P.w = 1; P.h = 1; P.filt_len = (int64_t)len;
    P.nblk = (int)((P.filt_len + kBlockBytes - 1) / kBlockBytes);
    P.nsub = 0;
    for (int b = 0; b < P.nblk; b++) {
        const int64_t l = std::min<int64_t>(kBlockBytes, P.filt_len - (int64_t)b * kBlockBytes);
        P.nsub += (int)((l + kSubBytes - 1) / kSubBytes);
    }
    P.png_bound = png_bound_of(P.filt_len, P.nblk);
    return 0;
}

int vcp_deflate(vcp_handle* h, const void* d_stream, uint64_t len, int bpp, int level, void* d_out, uint64_t cap, uint64_t* out_len) {
    if (!h || !d_out || !out_len) return fail(VCP_EINVAL, "bad arguments");
    LOCK_HANDLE(h);
    CU(cudaSetDevice(h->device));
    Lane& L = h->lane[0]; (void)L;
    std::vector<PagePlan> g(1);
    int rc = stream_plan(d_stream, len, bpp, g[0]); if (rc) return rc;
    vcp_opts o = {}; o.compress_level = level; o.src_device = 1;
    GroupOut go;
    rc = run_group(h, g, o, RUN_STREAM, go); if (rc) return rc;
    if (go.png_len[0] > cap) return fail(VCP_ESIZE, "output buffer too small: need %llu", (unsigned long long)go.png_len[0]);
    CU(cudaMemcpyAsync(d_out, go.d_png + go.png_off[0], go.png_len[0], cudaMemcpyDeviceToDevice, L.stream));
    CU(cudaStreamSynchronize(L.stream));
    *out_len = go.png_len[0];
    return 0;
}

int vcp_lz_sub_bytes(void) { return kSubBytes; }

int vcp_lz_tokens(vcp_handle* h, const void* d_stream, uint64_t len, int bpp, uint32_t* d_tokens, uint32_t* sub_ntok_host, uint32_t* sub_hist_host) {
    if (!h || !d_tokens || !sub_ntok_host) return fail(VCP_EINVAL, "bad arguments");
    LOCK_HANDLE(h);
    CU(cudaSetDevice(h->device));
    Lane& L = h->lane[0]; (void)L;
    std::vector<PagePlan> g(1);
    int rc = stream_plan(d_stream, len, bpp, g[0]); if (rc) return rc;
    vcp_opts o = {}; o.compress_level = 6; o.src_device = 1;
    GroupOut go;
    rc = run_group(h, g, o, RUN_LZ_ONLY, go); if (rc) return rc;
    // tokens are indexed by byte offset from the filtered region base; the stream sits kStreamPad after it
    CU(cudaMemcpyAsync(d_tokens, go.d_tokens + kStreamPad, (size_t)len * 4, cudaMemcpyDeviceToDevice, L.stream));
    CU(cudaMemcpyAsync(sub_ntok_host, go.d_sub_ntok, (size_t)go.nsub * 4, cudaMemcpyDeviceToHost, L.stream));
    if (sub_hist_host) CU(cudaMemcpyAsync(sub_hist_host, go.d_sub_hist, (size_t)go.nsub * kHistSize * 4, cudaMemcpyDeviceToHost, L.stream));
    CU(cudaStreamSynchronize(L.stream));
    return 0;
}

int vcp_adler32(vcp_handle* h, const void* d_data, uint64_t len, uint32_t* out) {
    if (!h || !out || (len && !d_data)) return fail(VCP_EINVAL, "bad arguments");
    LOCK_HANDLE(h);
    CU(cudaSetDevice(h->device));
    Lane& L = h->lane[0]; (void)L;
    const size_t nseg = (size_t)((len + 4095) / 4096);
    int rc = ensure_arena(L, 1024 + nseg * 8 + 512); if (rc) return rc;
    rc = ensure_meta(L, 4096); if (rc) return rc;
    uint32_t* d_out = reinterpret_cast<uint32_t*>(L.arena);
    launch_adler_flat((const uint8_t*)d_data, len, reinterpret_cast<uint32_t*>(L.arena + 1024), d_out, L.stream);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(L.meta, d_out, 4, cudaMemcpyDeviceToHost, L.stream));
    CU(cudaStreamSynchronize(L.stream));
    *out = *reinterpret_cast<uint32_t*>(L.meta);
    return 0;
}

int vcp_crc32(vcp_handle* h, const void* d_data, uint64_t len, uint32_t* out) {
    if (!h || !out || (len && !d_data)) return fail(VCP_EINVAL, "bad arguments");
    LOCK_HANDLE(h);
    CU(cudaSetDevice(h->device));
    Lane& L = h->lane[0]; (void)L;
    int rc = ensure_arena(L, 4096); if (rc) return rc;
    rc = ensure_meta(L, 4096); if (rc) return rc;
    uint32_t* d_out = reinterpret_cast<uint32_t*>(L.arena);
    launch_crc_flat((const uint8_t*)d_data, len, d_out, L.stream);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(L.meta, d_out, 4, cudaMemcpyDeviceToHost, L.stream));
    CU(cudaStreamSynchronize(L.stream));
    *out = *reinterpret_cast<uint32_t*>(L.meta);
    return 0;
}

int vcp_base64(vcp_handle* h, const void* d_src, uint64_t len, void* d_dst) {
    if (!h || (len && (!d_src || !d_dst))) return fail(VCP_EINVAL, "bad arguments");
    if (((uintptr_t)d_src & 3) || ((uintptr_t)d_dst & 15)) return fail(VCP_EINVAL, "vcp_base64 needs a 4-byte aligned source and a 16-byte aligned destination");
    LOCK_HANDLE(h);
    CU(cudaSetDevice(h->device));
    Lane& L = h->lane[0]; (void)L;
    launch_base64_flat((const uint8_t*)d_src, len, (uint8_t*)d_dst, L.stream);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(L.stream));
    return 0;
}

}  // extern "C"
