// esd.cu -- host side of libesd.so: the C ABI declared in include/esd.h.
//
// Replaces the scoring loop that /root/reference's scene task is specified to run
// (ModelManager.detect_scenes, ml-service/src/services/model_manager.py:715-835 ->
// PySceneDetect SceneManager/ContentDetector/AdaptiveDetector/HistogramDetector).
// No CPU fallback: every entry point needs a CUDA device.
#include "../../include/esd.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <dlfcn.h>
#include <string.h>
#include <stdlib.h>
#include <time.h>
#include <sched.h>
#include <memory>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <type_traits>
#include <map>
#include <string>
#include <vector>

#include "esd_kernels.cuh"
#include "host_tables.h"
#include "guard_alloc.h"
#include "ingest_gather.h"

using namespace esd;

namespace {

thread_local std::string g_create_error;

struct UnitPlan {
    Unit* d_units = nullptr;
    int* d_begin = nullptr;
    int grid = 0;
    int n_units = 0;
};

struct IngestSlot {
    uint8_t* h_pinned = nullptr;
    size_t h_pinned_bytes = 0;
    uint8_t* d_rows = nullptr;
    cudaEvent_t copied = nullptr;
    cudaEvent_t consumed = nullptr;
    bool in_flight = false;
};

// Host-side tap gather workers: ONE pool per process, shared by every context (node-level budget).  A context per GPU
// each with its own pool oversubscribed the host (8 ranks x 16 threads on 32 cores); here the pool
// has at most `budget()` threads -- the CPUs this process may run on, divided by the ranks sharing the node
// (LOCAL_WORLD_SIZE under torchrun) -- and concurrent parallel_for calls of several contexts (one feeding thread per
// GPU, eioku_b200.multi) share them: jobs queue up, workers drain the front job block by block, the submitting thread helps.
class SharedGatherPool {
   public:
    static SharedGatherPool& instance() {
        static SharedGatherPool* p = new SharedGatherPool();  // leaked on purpose: worker threads may outlive static destructors
        return *p;
    }
    static int budget() {
        int cpus = (int)std::thread::hardware_concurrency();
        cpu_set_t set;
        if (sched_getaffinity(0, sizeof set, &set) == 0) cpus = CPU_COUNT(&set);
        int ranks = 1;
        if (const char* e = getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, atoi(e));
        if (const char* e = getenv("ESD_NODE_RANKS")) ranks = std::max(1, atoi(e));
        return std::max(1, cpus / ranks);
    }
    // at least `n` workers (capped at the budget); returns the pool size
    int ensure(int n) {
        std::lock_guard<std::mutex> lk(m_);
        const int want = std::min(std::max(n, 1), budget());
        while ((int)workers_.size() < want) workers_.emplace_back([this] { run(); });
        return (int)workers_.size();
    }
    int size() {
        std::lock_guard<std::mutex> lk(m_);
        return (int)workers_.size();
    }
    // fn(lo, hi) over [0, n_items) in blocks; returns when every block has run.  `max_workers` bounds how many pool threads
    // may work on THIS job at once (a context that asked for fewer threads than the pool has keeps to its share).
    void parallel_for(int64_t n_items, int64_t block, int max_workers, const std::function<void(int64_t, int64_t)>& fn) {
        auto job = std::make_shared<Job>();
        job->fn = &fn;
        job->n = n_items;
        job->block = std::max<int64_t>(1, block);
        job->remaining.store(n_items);
        job->max_workers = std::max(1, max_workers);
        {
            std::lock_guard<std::mutex> lk(m_);
            queue_.push_back(job);
        }
        cv_.notify_all();
        work_on(*job, true);  // the submitter helps instead of sleeping
        std::unique_lock<std::mutex> lk(job->m);
        job->done.wait(lk, [&] { return job->remaining.load() <= 0; });
        {
            std::lock_guard<std::mutex> lk2(m_);
            queue_.erase(std::remove(queue_.begin(), queue_.end(), job), queue_.end());
        }
    }

   private:
    struct Job {
        const std::function<void(int64_t, int64_t)>* fn = nullptr;
        int64_t n = 0, block = 1;
        std::atomic<int64_t> next{0}, remaining{0};
        std::atomic<int> active{0};
        int max_workers = 1;
        std::mutex m;
        std::condition_variable done;
    };
    void work_on(Job& j, bool submitter) {
        if (!submitter && j.active.fetch_add(1) >= j.max_workers - 1) {  // the submitter counts as one worker of its job
            j.active.fetch_sub(1);
            return;
        }
        for (;;) {
            const int64_t lo = j.next.fetch_add(j.block);
            if (lo >= j.n) break;
            const int64_t hi = std::min(j.n, lo + j.block);
            (*j.fn)(lo, hi);
            if (j.remaining.fetch_sub(hi - lo) - (hi - lo) <= 0) {
                std::lock_guard<std::mutex> lk(j.m);
                j.done.notify_all();
            }
        }
        if (!submitter) j.active.fetch_sub(1);
    }
    void run() {
        for (;;) {
            std::shared_ptr<Job> job;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] {
                    for (auto& q : queue_)
                        if (q->next.load() < q->n && q->active.load() < q->max_workers - 1) { job = q; return true; }
                    return false;
                });
            }
            work_on(*job, false);
        }
    }
    std::mutex m_;
    std::condition_variable cv_;
    std::vector<std::thread> workers_;
    std::vector<std::shared_ptr<Job>> queue_;
};

// Kernel launcher of push_common.  DIRECT launches on the stream; the per-frame plugin path (esd_process_frame_host) is
// launch-bound (three to seven tiny kernels per call), so there the same call sequence BUILDs a CUDA graph once (one
// kernel node per launch, chained) and afterwards only UPDATEs the nodes' parameters before a single cudaGraphLaunch.
struct KernelGraph {
    enum Mode { DIRECT = 0, BUILD = 1, UPDATE = 2 };
    int mode = DIRECT;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    std::vector<cudaGraphNode_t> nodes;
    size_t cursor = 0;
    void destroy() {
        if (exec) cudaGraphExecDestroy(exec);
        if (graph) cudaGraphDestroy(graph);
        exec = nullptr; graph = nullptr; nodes.clear(); cursor = 0; mode = DIRECT;
    }
};

template <typename... P>
cudaError_t klaunch(KernelGraph& g, cudaStream_t st, void (*kernel)(P...), dim3 grid, dim3 block, size_t smem,
                    typename std::decay<P>::type... args) {
    void* argv[] = {(void*)&args...};
    if (g.mode == KernelGraph::DIRECT) return cudaLaunchKernel((const void*)kernel, grid, block, argv, smem, st);
    cudaKernelNodeParams kp{};
    kp.func = (void*)kernel;
    kp.gridDim = grid;
    kp.blockDim = block;
    kp.sharedMemBytes = (unsigned)smem;
    kp.kernelParams = argv;
    if (g.mode == KernelGraph::BUILD) {
        cudaGraphNode_t node;
        const cudaGraphNode_t* dep = g.nodes.empty() ? nullptr : &g.nodes.back();
        cudaError_t e = cudaGraphAddKernelNode(&node, g.graph, dep, g.nodes.empty() ? 0 : 1, &kp);
        if (e == cudaSuccess) g.nodes.push_back(node);
        return e;
    }
    if (g.cursor >= g.nodes.size()) return cudaErrorInvalidValue;  // topology changed: caller rebuilds
    return cudaGraphExecKernelNodeSetParams(g.exec, g.nodes[g.cursor++], &kp);
}

}  // namespace

struct esd_ctx {
    esd_config cfg;
    int device = 0;
    int num_sms = 0;
    std::string err;

    // geometry
    int dst_w = 0, dst_h = 0, row_bytes = 0;
    int alg_row_bytes = 0;  // 32-byte sectors of a row that contain a horizontal tap, in bytes
    int64_t alg_frame_bytes = 0;  // sum over the touched rows (NV12: Y rows and UV rows have different sector sets)
    bool nv12 = false;      // the fused kernel's YUV 4:2:0 variant (NV12 and I420 contexts)
    bool i420 = false;      // planar chroma: the touched U / V rows are interleaved into NV12 rows first (i420_interleave_kernel)
    int* d_uv_src_rows = nullptr;   // I420: chroma row of each touched UV row, in `touched` order
    uint8_t* d_uvpack = nullptr;    // I420 device pushes: [n][touched UV rows][row_bytes] interleaved chroma (grow-only)
    size_t uvpack_bytes = 0;
    bool extras = false;    // the fused kernel also emits sum(B+G+R) / the V plane / the gray plane (kernel template EXTRAS)
    int n_touched_y = 0;    // NV12: the first n_touched_y entries of `touched` are Y rows, the rest UV rows (as H + row)
    bool resize = false;
    int pxt = 1;
    int rows_per_group = 0, n_groups = 0, stages = 0, rowbuf = 0, stage_bytes = 0, rows_per_stage = 1;
    int ctas_per_sm = 0;
    int lane_stride = 1, lane_stride_taps = 1;  // consumer lane -> column stride (host_tables.h: choose_lane_stride), full rows / gathered taps
    size_t smem_bytes = 0;
    bool need_content = false, need_hist = false, need_edges = false, need_hash = false;
    int edge_ksize = 0, edge_words = 0;
    HashParams hparams{};          // perceptual hash geometry + device tables
    int hash_words = 0;
    size_t hash_smem = 0;
    void (*hash_fn)(const uint8_t*, int, HashParams, uint8_t*, uint32_t*) = nullptr;
    int* d_hash_itab = nullptr;    // area-resize tables: x begin, y begin, x src, y src (one allocation)
    float* d_hash_wtab = nullptr;  // x weights, y weights
    double* d_hash_C = nullptr;    // [hash_size][S] DCT rows
    std::vector<int32_t> touched;  // ascending source rows
    ScoreWeights wc{}, wa{};
    DecisionParams dparams{};

    // device tables
    YRow* d_yrows = nullptr;
    uint2* d_xtab = nullptr;
    int* d_sdiv = nullptr;
    int* d_hdiv = nullptr;

    // per-video state
    int64_t first_frame = 0;
    int64_t n_frames = 0;
    bool started = false;
    int prev_parity = 0;
    uint32_t* d_prev[2] = {nullptr, nullptr};
    DecisionState* d_state = nullptr;
    long long* d_cuts = nullptr;  // [5][max_cuts]
    int64_t max_cuts = 0;

    // growable per-frame score arrays
    int64_t cap = 0;
    unsigned long long* d_sums3 = nullptr;
    double* d_cv = nullptr;
    double* d_av = nullptr;
    double* d_ratio = nullptr;
    double* d_hdiff = nullptr;
    double* d_avg = nullptr;       // ThresholdDetector average_rgb
    uint32_t* d_edge_counts = nullptr;  // differing edge pixels vs the previous frame
    uint32_t* d_counts = nullptr;  // [cap][bins]
    uint32_t* d_hash = nullptr;    // [cap][hash_words] perceptual hash bits
    double* d_hdist = nullptr;     // HashDetector hash_dist (normalised Hamming distance to the previous frame)
    float* d_hmargin = nullptr;    // HashDetector: smallest |DCT coefficient - median| per frame (stability of its weakest bit)
    uint8_t* d_slab = nullptr;     // backing allocation of the six arrays above

    // per-batch scratch
    int64_t part_cap_frames = 0;
    uint4* d_part[2] = {nullptr, nullptr};          // double-buffered: the tail of push k overlaps push k+1
    uint16_t* d_hist_part[2] = {nullptr, nullptr};
    uint8_t* d_vplane[2] = {nullptr, nullptr};      // V planes of the batch (edge detector input)
    uint8_t* d_gplane[2] = {nullptr, nullptr};      // BGR2GRAY planes of the batch (hash detector input)
    uint8_t* d_hash_small = nullptr;                // [n][S * S] INTER_AREA thumbnails of the batch (test hook)
    int64_t last_batch_base = 0, last_batch_n = 0;
    uint32_t* d_edge_bits = nullptr;                // [n][edge_words] dilated edge bitmaps of the batch
    uint32_t* d_edge_prev = nullptr;                // [edge_words] bitmap of the last frame of the previous batch
    int part_buf = 0;
    std::map<int64_t, UnitPlan> plans;

    // streams / timing
    cudaStream_t last_stream = nullptr;
    bool have_last_stream = false;
    cudaEvent_t order_event = nullptr;
    // finalize + decision kernels run on a library-owned stream so they overlap the next batch's fused kernel
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_fused = nullptr, ev_join = nullptr;
    cudaEvent_t ev_fin[2] = {nullptr, nullptr};
    bool fin_recorded[2] = {false, false};
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timing_events;
    int64_t launches = 0;

    // per-frame plugin path (esd_process_frame_host): one pinned + one device frame, results through pinned memory
    uint8_t* pf_h = nullptr;           // pinned frame, read by the fused kernel's TMA loads straight over PCIe
    long long* pf_mailbox = nullptr;   // pinned, host-mapped: n_cuts[5] + overflow flag written by decide_kernel
    cudaStream_t pf_stream = nullptr;
    KernelGraph kg;                    // DIRECT except inside esd_process_frame_host
    int64_t pf_ticket = 0;             // sequence number decide_kernel posts into the mailbox when a detector's pass is done

    // stand-alone decision pass (esd_decide_arrays / esd_decide_device): persistent, grow-only scratch -- no allocation,
    // no legacy-stream launch and no device-wide synchronisation per call
    cudaStream_t dec_stream = nullptr;
    double* dec_d_scores = nullptr;      // [dec_cap] staged host scores
    double* dec_d_ratio = nullptr;       // [dec_cap] adaptive ratios of the pass
    int64_t dec_cap = 0;
    DecisionState* dec_d_state = nullptr;
    uint8_t* dec_h = nullptr;            // pinned, host-mapped: [16 x i64 mailbox][max_cuts x i64 cuts][dec_h_scores x f64 staging]
    int64_t dec_h_scores = 0;
    int64_t dec_ticket = 0;
    bool poisoned = false;               // a push failed after its fused kernel was enqueued: state is undefined until esd_reset

    // ingest
    std::vector<IngestSlot> ring;
    int frames_per_slot = 0;
    int next_slot = 0;
    cudaStream_t copy_stream = nullptr, compute_stream = nullptr;
    int64_t h2d_bytes = 0, h2d_copies = 0;
    // tap gather: host threads copy only the 6 bytes (two BGR taps) each destination column reads from a touched row
    int gather_threads = 0;              // > 0: gather enabled, at most this many threads of the process-wide pool per push
    int tap_row_bytes = 0;               // 6 * dst_w rounded up to 16
    uint2* d_xtab_taps = nullptr;        // x table for the tap-compact layout {6 d, a0 | a1 << 16}
    std::vector<int> tap_src_off;        // byte offset of tap 0 in a source row, per destination column
};

namespace {

int fail(esd_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    else g_create_error = buf;
    return code;
}

#define CU(c, call)                                                                                   \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail((c), ESD_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                        __LINE__);                                                                    \
    } while (0)

int sync_all(esd_ctx* c) {
    if (c->copy_stream) CU(c, cudaStreamSynchronize(c->copy_stream));
    if (c->have_last_stream) CU(c, cudaStreamSynchronize(c->last_stream));
    if (c->aux_stream) CU(c, cudaStreamSynchronize(c->aux_stream));
    return ESD_OK;
}

// Python round(): half to even on the double quotient
long py_round(double x) { return lrint(x); }

// the kernel instance for a (content, hist, aligned, nv12, extras) combination; NV12 exists only for resizing contexts
template <bool RESIZE, int PXT, bool C, bool H>
void (*fused_fn_ch(bool aligned, bool nv12, bool extras))(const FusedParams) {
    if constexpr (RESIZE && PXT <= 4) {
        if (nv12) {
            if (extras) return aligned ? fused_score_kernel<RESIZE, PXT, C, H, true, true, true> : fused_score_kernel<RESIZE, PXT, C, H, false, true, true>;
            return aligned ? fused_score_kernel<RESIZE, PXT, C, H, true, true, false> : fused_score_kernel<RESIZE, PXT, C, H, false, true, false>;
        }
    }
    if (extras) return aligned ? fused_score_kernel<RESIZE, PXT, C, H, true, false, true> : fused_score_kernel<RESIZE, PXT, C, H, false, false, true>;
    return aligned ? fused_score_kernel<RESIZE, PXT, C, H, true, false, false> : fused_score_kernel<RESIZE, PXT, C, H, false, false, false>;
}

template <bool RESIZE, int PXT>
void (*fused_fn(bool content, bool hist, bool aligned, bool nv12, bool extras))(const FusedParams) {
    if (content && hist) return fused_fn_ch<RESIZE, PXT, true, true>(aligned, nv12, extras);
    if (content) return fused_fn_ch<RESIZE, PXT, true, false>(aligned, nv12, extras);
    return fused_fn_ch<RESIZE, PXT, false, true>(aligned, nv12, extras);
}

template <bool RESIZE, int PXT>
cudaError_t launch_fused_rp(bool content, bool hist, bool aligned, bool nv12, bool extras, const FusedParams& p, int grid, size_t smem,
                            cudaStream_t st, KernelGraph& kg) {
    return klaunch(kg, st, fused_fn<RESIZE, PXT>(content, hist, aligned, nv12, extras), dim3(grid), dim3(kThreads), smem, p);
}

template <bool RESIZE, int PXT>
cudaError_t occupancy_rp(bool content, bool hist, bool nv12, bool extras, size_t smem, int* out) {
    for (int al = 0; al < 2; ++al) {
        cudaError_t e = cudaFuncSetAttribute(fused_fn<RESIZE, PXT>(content, hist, al != 0, nv12, extras),
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, fused_fn<RESIZE, PXT>(content, hist, true, nv12, extras), kThreads, smem);
}

#define ESD_DISPATCH(FN, resize, pxt, ...)                                           \
    ((resize) ? ((pxt) == 1 ? FN<true, 1>(__VA_ARGS__)                               \
                 : (pxt) == 2 ? FN<true, 2>(__VA_ARGS__)                             \
                              : FN<true, 4>(__VA_ARGS__))                            \
              : ((pxt) == 1 ? FN<false, 1>(__VA_ARGS__)                              \
                 : (pxt) == 2 ? FN<false, 2>(__VA_ARGS__)                            \
                 : (pxt) == 4 ? FN<false, 4>(__VA_ARGS__)                            \
                 : (pxt) == 8 ? FN<false, 8>(__VA_ARGS__)                            \
                              : FN<false, 16>(__VA_ARGS__)))

size_t fused_smem_bytes(const esd_ctx* c, int R, int stages) {
    size_t off = 4864;  // barriers + meta + per-row meta + sdiv/hdiv + hist (see kernel carve-up)
    if (c->need_content) off += (size_t)R * c->pxt * kConsumers * 4;
    off = (off + 127) & ~(size_t)127;
    return off + (size_t)stages * c->stage_bytes;
}

void free_plans(esd_ctx* c) {
    for (auto& kv : c->plans) {
        esdguard::gfree(kv.second.d_units);
        esdguard::gfree(kv.second.d_begin);
    }
    c->plans.clear();
}

// Work decomposition.  A unit is (row group, frame range); units that do not start at frame 0
// re-read one halo frame to rebuild the previous HSV, so fewer/longer units are cheaper.
int build_plan(esd_ctx* c, int64_t n, UnitPlan* out) {
    std::vector<Unit> units;
    std::vector<int> begin;
    // reserved2 > 0 caps the grid: a context that shares the GPU with other kernels (e.g. a decoder whose blocks own whole SMs'
    // shared memory) must not wait for every SM to become free for a batch that 16 CTAs finish in a few hundred microseconds
    int64_t max_ctas = (int64_t)c->num_sms * c->ctas_per_sm;
    if (c->cfg.reserved2 > 0) max_ctas = std::min<int64_t>(max_ctas, c->cfg.reserved2);
    const int grid = build_unit_plan(c->n_groups, n, max_ctas,
                                     c->cfg.split_mode ? c->cfg.split_mode : ESD_SPLIT_STRIPS, units, begin);
    out->grid = grid;
    out->n_units = (int)units.size();
    CU(c, esdguard::gmalloc(&out->d_units, sizeof(Unit) * std::max<size_t>(1, units.size())));
    CU(c, esdguard::gmalloc(&out->d_begin, sizeof(int) * begin.size()));
    CU(c, cudaMemcpy(out->d_units, units.data(), sizeof(Unit) * units.size(), cudaMemcpyHostToDevice));
    CU(c, cudaMemcpy(out->d_begin, begin.data(), sizeof(int) * begin.size(), cudaMemcpyHostToDevice));
    return ESD_OK;
}

// Per-frame score arrays live in ONE device slab (a single cudaMalloc per growth: cudaMalloc/cudaFree cost
// milliseconds each on a busy device and sit on the first push's latency).
int ensure_capacity(esd_ctx* c, int64_t frames_needed) {
    if (frames_needed <= c->cap) return ESD_OK;
    const int64_t ncap = std::max<int64_t>(c->cap ? c->cap * 2 : std::max<int64_t>(4096, c->cfg.initial_capacity), frames_needed);
    { int rc0 = sync_all(c); if (rc0) return rc0; }
    const int64_t bins = c->need_hist ? c->cfg.hist_bins : 0;
    const int64_t hwords = c->need_hash ? c->hash_words : 0;
    const size_t bytes = (size_t)ncap * (3 * sizeof(unsigned long long) + 6 * sizeof(double) + (bins + 2 + hwords) * sizeof(uint32_t));
    uint8_t* slab = nullptr;
    CU(c, esdguard::gmalloc(&slab, bytes));
    uint8_t* q = slab;
    auto carve = [&](auto** dst, int64_t per_frame) {
        using T = typename std::remove_pointer<typename std::remove_pointer<decltype(dst)>::type>::type;
        T* np = reinterpret_cast<T*>(q);
        q += sizeof(T) * per_frame * ncap;
        return np;
    };
    const int64_t used = c->n_frames;
    auto* n_sums = carve(&c->d_sums3, 3);
    auto* n_cv = carve(&c->d_cv, 1);
    auto* n_av = carve(&c->d_av, 1);
    auto* n_ratio = carve(&c->d_ratio, 1);
    auto* n_hdiff = carve(&c->d_hdiff, 1);
    auto* n_avg = carve(&c->d_avg, 1);
    auto* n_hdist = carve(&c->d_hdist, 1);
    auto* n_edge = carve(&c->d_edge_counts, 1);
    auto* n_counts = carve(&c->d_counts, bins);
    auto* n_hash = carve(&c->d_hash, hwords);
    auto* n_hmargin = carve(&c->d_hmargin, 1);
    if (used > 0) {
        CU(c, cudaMemcpy(n_sums, c->d_sums3, sizeof(unsigned long long) * 3 * used, cudaMemcpyDeviceToDevice));
        CU(c, cudaMemcpy(n_cv, c->d_cv, sizeof(double) * used, cudaMemcpyDeviceToDevice));
        CU(c, cudaMemcpy(n_av, c->d_av, sizeof(double) * used, cudaMemcpyDeviceToDevice));
        CU(c, cudaMemcpy(n_ratio, c->d_ratio, sizeof(double) * used, cudaMemcpyDeviceToDevice));
        CU(c, cudaMemcpy(n_hdiff, c->d_hdiff, sizeof(double) * used, cudaMemcpyDeviceToDevice));
        CU(c, cudaMemcpy(n_avg, c->d_avg, sizeof(double) * used, cudaMemcpyDeviceToDevice));
        CU(c, cudaMemcpy(n_edge, c->d_edge_counts, sizeof(uint32_t) * used, cudaMemcpyDeviceToDevice));
        if (bins) CU(c, cudaMemcpy(n_counts, c->d_counts, sizeof(uint32_t) * bins * used, cudaMemcpyDeviceToDevice));
        CU(c, cudaMemcpy(n_hdist, c->d_hdist, sizeof(double) * used, cudaMemcpyDeviceToDevice));
        if (hwords) CU(c, cudaMemcpy(n_hash, c->d_hash, sizeof(uint32_t) * hwords * used, cudaMemcpyDeviceToDevice));
        CU(c, cudaMemcpy(n_hmargin, c->d_hmargin, sizeof(float) * used, cudaMemcpyDeviceToDevice));
    }
    esdguard::gfree(c->d_slab);
    c->d_slab = slab;
    c->d_sums3 = n_sums; c->d_cv = n_cv; c->d_av = n_av; c->d_ratio = n_ratio; c->d_hdiff = n_hdiff; c->d_avg = n_avg; c->d_edge_counts = n_edge;
    c->d_counts = bins ? n_counts : nullptr;
    c->d_hdist = n_hdist;
    c->d_hash = hwords ? n_hash : nullptr;
    c->d_hmargin = n_hmargin;
    // ratios not yet computed read back as NaN
    fill_nan_kernel<<<(unsigned)((ncap - used + 255) / 256), 256, 0, c->aux_stream>>>(c->d_ratio + used, ncap - used);
    CU(c, cudaGetLastError());
    CU(c, cudaStreamSynchronize(c->aux_stream));
    c->cap = ncap;
    return ESD_OK;
}

int ensure_scratch(esd_ctx* c, int64_t n) {
    if (n <= c->part_cap_frames) return ESD_OK;
    { int rc0 = sync_all(c); if (rc0) return rc0; }
    c->part_cap_frames = 0;
    for (int b = 0; b < 2; ++b) {
        esdguard::gfree(c->d_part[b]);
        esdguard::gfree(c->d_hist_part[b]);
        esdguard::gfree(c->d_vplane[b]);
        esdguard::gfree(c->d_gplane[b]);
        c->d_gplane[b] = nullptr;
        c->d_part[b] = nullptr;
        c->d_hist_part[b] = nullptr;
        c->d_vplane[b] = nullptr;
        c->fin_recorded[b] = false;
        if (c->need_edges) CU(c, esdguard::gmalloc(&c->d_vplane[b], (size_t)n * c->dst_w * c->dst_h));
        if (c->need_hash) CU(c, esdguard::gmalloc(&c->d_gplane[b], (size_t)n * c->dst_w * c->dst_h + 16));  // + slack: word loads
        if (c->need_content) CU(c, esdguard::gmalloc(&c->d_part[b], sizeof(uint4) * n * c->n_groups * kConsumerWarps));
        if (c->need_hist) CU(c, esdguard::gmalloc(&c->d_hist_part[b], sizeof(uint16_t) * n * c->n_groups * c->cfg.hist_bins));
    }
    if (c->need_edges) {
        esdguard::gfree(c->d_edge_bits);
        c->d_edge_bits = nullptr;
        CU(c, esdguard::gmalloc(&c->d_edge_bits, sizeof(uint32_t) * n * c->edge_words));
    }
    if (c->need_hash) {
        esdguard::gfree(c->d_hash_small);
        c->d_hash_small = nullptr;
        CU(c, esdguard::gmalloc(&c->d_hash_small, (size_t)n * c->hparams.S * c->hparams.S));
    }
    c->part_cap_frames = n;
    return ESD_OK;
}

int reset_video_state(esd_ctx* c) {
    c->poisoned = false;
    c->first_frame = 0;
    c->n_frames = 0;
    c->started = false;
    c->prev_parity = 0;
    c->h2d_bytes = 0;
    c->h2d_copies = 0;
    c->last_batch_base = c->last_batch_n = 0;
    if (c->pf_mailbox) memset(c->pf_mailbox, 0, 8 * sizeof(long long));  // counts and overflow; the tickets keep increasing
    // on the library's own stream, waiting for this context only: a reset must not stall the other contexts of the device
    // (several decoder sessions per GPU reset their context once per video)
    CU(c, cudaMemsetAsync(c->d_state, 0, sizeof(DecisionState), c->aux_stream));
    if (c->cap) {
        fill_nan_kernel<<<(unsigned)((c->cap + 255) / 256), 256, 0, c->aux_stream>>>(c->d_ratio, c->cap);
        CU(c, cudaGetLastError());
    }
    CU(c, cudaStreamSynchronize(c->aux_stream));
    return ESD_OK;
}

// make `st` wait for everything previously enqueued for this ctx on another stream
int order_after_last(esd_ctx* c, cudaStream_t st) {
    if (c->have_last_stream && c->last_stream != st) {
        CU(c, cudaEventRecord(c->order_event, c->last_stream));
        CU(c, cudaStreamWaitEvent(st, c->order_event, 0));
    }
    c->last_stream = st;
    c->have_last_stream = true;
    return ESD_OK;
}


// Caller-supplied frame pointers are validated before a kernel touches them: an out-of-bounds TMA read is a sticky
// CUDA fault.  Pageable host memory is rejected; for device allocations the driver's cuMemGetAddressRange (resolved
// lazily, so the library still loads on machines without a driver) bounds the [first byte, last byte] span.
int validate_device_span(esd_ctx* c, const uint8_t* ptr, size_t span_bytes, const char* what) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, ptr) != cudaSuccess) {
        cudaGetLastError();
        return fail(c, ESD_ERR_INVALID, "%s: pointer is not known to CUDA", what);
    }
    if (attr.type == cudaMemoryTypeUnregistered)
        return fail(c, ESD_ERR_INVALID, "%s: pageable host memory is not device-accessible (use esd_ingest_push_host)", what);
    if (attr.type != cudaMemoryTypeDevice) return ESD_OK;  // pinned / managed memory: reachable, extent unknown
    typedef int (*get_range_fn)(unsigned long long*, size_t*, unsigned long long);
    static get_range_fn get_range = [] {
        void* h = dlopen("libcuda.so.1", RTLD_LAZY | RTLD_LOCAL);
        return h ? (get_range_fn)dlsym(h, "cuMemGetAddressRange_v2") : (get_range_fn) nullptr;
    }();
    if (!get_range) return ESD_OK;
    unsigned long long base = 0;
    size_t size = 0;
    if (get_range(&base, &size, (unsigned long long)(uintptr_t)ptr) != 0) return ESD_OK;
    const unsigned long long end = (unsigned long long)(uintptr_t)ptr + span_bytes;
    if ((unsigned long long)(uintptr_t)ptr < base || end > base + size)
        return fail(c, ESD_ERR_INVALID, "%s: frames span %zu bytes but the allocation ends %lld bytes earlier", what, span_bytes,
                    (long long)(end - (base + size)));
    return ESD_OK;
}

struct TraceTimer {
    bool on = getenv("ESD_TRACE") != nullptr;
    double t0 = now();
    static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
    void lap(const char* what) {
        if (!on) return;
        const double t = now();
        fprintf(stderr, "[esd trace] %-18s %.3f ms\n", what, t - t0);
        t0 = t;
    }
};

enum Layout { LAYOUT_FULL = 0, LAYOUT_ROWS = 1, LAYOUT_TAPS = 2 };
// I420 device pushes: the touched chroma rows, interleaved, in a buffer of their own ([n][touched UV rows] rows of `row_stride`)
struct UvPacked { int64_t frame_stride, row_stride; };

// inline_tail: run the finalize/decision tail on `st` itself (per-frame path: nothing to overlap with, and the
// cross-stream event round trip would cost more than the tail)
int push_common(esd_ctx* c, const uint8_t* d_src, int64_t n, int64_t frame_stride, int64_t row_stride, int layout,
                int64_t first_frame_num, cudaStream_t st, bool inline_tail = false, long long* mailbox = nullptr,
                const uint8_t* d_uv = nullptr, const UvPacked* uvp = nullptr) {
    const bool compact = layout != LAYOUT_FULL;
    TraceTimer tr;
    if (!d_src || n <= 0) return fail(c, ESD_ERR_INVALID, "push: null frames or n <= 0");
    if (n > 0x7fffff00LL / std::max(1, c->n_groups)) return fail(c, ESD_ERR_INVALID, "push: batch too large (%lld frames)", (long long)n);
    if (c->poisoned) return fail(c, ESD_ERR_STATE, "push: an earlier push failed half-way; call esd_reset");
    CU(c, cudaSetDevice(c->device));
    // Per-video state (started / first_frame, scratch and previous-frame parities, n_frames) is committed only once the fused
    // kernel is enqueued: a push that fails earlier leaves the ctx untouched and may be retried; one that fails later
    // (tail launches, graph node updates) poisons the ctx until esd_reset instead of silently diffing against a half-made batch.
    struct PoisonGuard {
        esd_ctx* c; bool armed = false, ok = false;
        ~PoisonGuard() { if (armed && !ok) c->poisoned = true; }
    } guard{c};
    if (c->started && first_frame_num != c->first_frame + c->n_frames) {
        return fail(c, ESD_ERR_STATE, "push: frame numbers must be sequential (expected %lld, got %lld)",
                    (long long)(c->first_frame + c->n_frames), (long long)first_frame_num);
    }
    int rc;
    if ((rc = ensure_capacity(c, c->n_frames + n))) return rc;
    tr.lap("ensure_capacity");
    if ((rc = ensure_scratch(c, n))) return rc;
    tr.lap("ensure_scratch");
    auto it = c->plans.find(n);
    if (it == c->plans.end()) {
        if (c->plans.size() > 64) free_plans(c);
        UnitPlan pl;
        if ((rc = build_plan(c, n, &pl))) return rc;
        it = c->plans.emplace(n, pl).first;
    }
    const UnitPlan& plan = it->second;
    tr.lap("plan");
    if ((rc = order_after_last(c, st))) return rc;

    const int64_t base = c->n_frames;  // index of the batch's first frame
    FusedParams p{};
    p.src = d_src;
    p.src_uv = d_uv ? d_uv : d_src + (int64_t)c->cfg.src_height * row_stride;  // NV12 default: UV plane right behind the Y plane
    p.frame_stride = frame_stride;
    p.row_stride = row_stride;
    p.uv_frame_stride = uvp ? uvp->frame_stride : frame_stride;
    p.uv_row_stride = uvp ? uvp->row_stride : row_stride;
    p.uv_packed_base = uvp ? c->n_touched_y : -1;
    p.compact = compact ? 1 : 0;
    p.n_frames = (int)n;
    p.dst_w = c->dst_w;
    p.dst_h = c->dst_h;
    p.row_bytes = layout == LAYOUT_TAPS ? c->tap_row_bytes : c->row_bytes;
    p.rows_per_group = c->rows_per_group;
    p.n_groups = c->n_groups;
    p.stages = c->stages;
    p.rows_per_stage = c->rows_per_stage;
    p.rowbuf = c->rowbuf;
    p.has_prev = base > 0 ? 1 : 0;
    p.bins = c->cfg.hist_bins;
    p.want_bgr = (c->cfg.detectors & ESD_DET_THRESHOLD) ? 1 : 0;
    p.lane_stride = layout == LAYOUT_TAPS ? c->lane_stride_taps : c->lane_stride;
    p.yrows = c->d_yrows;
    p.xtab = layout == LAYOUT_TAPS ? c->d_xtab_taps : c->d_xtab;
    p.sdiv = c->d_sdiv;
    p.hdiv = c->d_hdiv;
    p.units = plan.d_units;
    p.cta_unit_begin = plan.d_begin;
    p.prev_in = c->d_prev[c->prev_parity];
    p.prev_out = c->d_prev[c->prev_parity ^ 1];
    const int buf = c->part_buf;
    p.part = c->d_part[buf];
    p.hist_part = c->d_hist_part[buf];
    p.vplane = c->need_edges ? c->d_vplane[buf] : nullptr;
    p.gplane = c->need_hash ? c->d_gplane[buf] : nullptr;
    // the fused kernel overwrites part[buf]: wait until the tail of the push that last used it is done
    // (the per-frame path is synchronous: nothing of it is ever in flight here)
    if (c->fin_recorded[buf] && !inline_tail) CU(c, cudaStreamWaitEvent(st, c->ev_fin[buf], 0));
    KernelGraph& kg = c->kg;

    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (c->timing) {
        CU(c, cudaEventCreate(&e0));
        CU(c, cudaEventCreate(&e1));
        CU(c, cudaEventRecord(e0, st));
    }
    const bool aligned = ((reinterpret_cast<uintptr_t>(d_src) | (c->nv12 ? reinterpret_cast<uintptr_t>(p.src_uv) : 0) |
                           (uintptr_t)frame_stride | (uintptr_t)row_stride | (uintptr_t)p.uv_frame_stride | (uintptr_t)p.uv_row_stride) & 15u) == 0;
    CU(c, ESD_DISPATCH(launch_fused_rp, c->resize, c->pxt, c->need_content, c->need_hist, aligned, c->nv12, c->extras, p, plan.grid, c->smem_bytes, st, kg));
    guard.armed = true;
    if (!c->started) {
        c->first_frame = first_frame_num;
        c->started = true;
    }
    c->part_buf ^= 1;
    c->prev_parity ^= 1;
    c->launches++;
    tr.lap("launch fused");
    if (c->timing) {
        CU(c, cudaEventRecord(e1, st));
        c->timing_events.emplace_back(e0, e1);
    }
    // tail (finalize + decision) on the library's own stream, ordered after this fused kernel
    cudaStream_t ts = inline_tail ? st : c->aux_stream;
    if (!inline_tail) {
        CU(c, cudaEventRecord(c->ev_fused, st));
        CU(c, cudaStreamWaitEvent(ts, c->ev_fused, 0));
    }

    const double npx = (double)c->dst_w * (double)c->dst_h;  // float(rows * cols)
    if (c->need_edges) {
        const size_t esmem = 2 * (size_t)(((c->dst_w * c->dst_h) + 31) & ~31) + 8 * (size_t)c->edge_words;
        CU(c, klaunch(kg, ts, edges_kernel, dim3((unsigned)n), dim3(kEdgeThreads), esmem, c->d_vplane[buf], c->dst_w, c->dst_h,
                      c->edge_ksize, c->d_edge_bits, c->edge_words));
        CU(c, klaunch(kg, ts, edge_delta_kernel, dim3((unsigned)n), dim3(256), 0, c->d_edge_bits, c->d_edge_prev, c->edge_words,
                      base > 0 ? 1 : 0, c->d_edge_counts + base));
        CU(c, cudaMemcpyAsync(c->d_edge_prev, c->d_edge_bits + (size_t)(n - 1) * c->edge_words, sizeof(uint32_t) * c->edge_words,
                              cudaMemcpyDeviceToDevice, ts));
        c->launches += 2;
    }
    if (c->need_content) {
        const int warps_per_block = 8;
        CU(c, klaunch(kg, ts, finalize_sums_kernel, dim3((unsigned)((n + warps_per_block - 1) / warps_per_block)),
                      dim3(warps_per_block * 32), 0, c->d_part[buf], (int)n, c->n_groups * kConsumerWarps, npx, c->wc, c->wa,
                      c->d_sums3 + 3 * base, c->d_cv + base, c->d_av + base, c->d_avg + base,
                      c->need_edges ? c->d_edge_counts + base : nullptr));
        c->launches++;
        if (c->cfg.detectors & ESD_DET_ADAPTIVE) {
            const int w = c->cfg.adaptive_window_width;
            const int64_t t0 = std::max<int64_t>(w, base - w), t1 = base + n - w;
            if (t1 > t0) {
                CU(c, klaunch(kg, ts, adaptive_ratio_kernel, dim3((unsigned)((t1 - t0 + 255) / 256)), dim3(256), 0, c->d_av,
                              c->d_ratio, (long long)t0, (long long)t1, w, c->cfg.adaptive_min_content_val));
                c->launches++;
            }
        }
    }
    if (c->need_hist) {
        const int bins = c->cfg.hist_bins;
        CU(c, klaunch(kg, ts, finalize_hist_counts_kernel, dim3((unsigned)n), dim3(256), 0, c->d_hist_part[buf], c->n_groups, bins,
                      c->d_counts + base * bins));
        CU(c, klaunch(kg, ts, hist_diff_kernel, dim3((unsigned)n), dim3(256), 0, c->d_counts + base * bins, bins, base > 0 ? 1 : 0,
                      c->d_hdiff + base));
        c->launches += 2;
    }
    if (c->need_hash) {
        HashParams hp = c->hparams;
        hp.margin_out = c->d_hmargin + base;
        const size_t hsmem = c->hash_smem;
        CU(c, klaunch(kg, ts, c->hash_fn, dim3((unsigned)n), dim3(kHashThreads), hsmem, c->d_gplane[buf], (int)n, hp, c->d_hash_small,
                      c->d_hash + base * c->hash_words));
        CU(c, klaunch(kg, ts, hash_dist_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, c->d_hash + base * c->hash_words,
                      c->hash_words, (int)n, base > 0 ? 1 : 0, (double)(hp.hs * hp.hs), c->d_hdist + base));
        c->launches += 2;
        c->last_batch_base = base;
        c->last_batch_n = n;
    }
    CU(c, klaunch(kg, ts, decide_kernel, dim3(5), dim3(kDecideThreads), 0, c->dparams, c->d_state, c->d_cuts, c->d_cv, c->d_av, c->d_ratio,
                  c->d_hdiff, c->d_avg, c->d_hdist, (long long)c->first_frame, (long long)base, (long long)(base + n), mailbox,
                  (long long)c->pf_ticket));
    if (!inline_tail) {
        CU(c, cudaEventRecord(c->ev_fin[buf], ts));
        c->fin_recorded[buf] = true;
    }
    c->launches++;
    c->n_frames += n;
    guard.ok = true;
    tr.lap("launch tail");
    return ESD_OK;
}

}  // namespace

// ======================================================================================= C ABI
extern "C" {

int esd_abi_version(void) { return ESD_ABI_VERSION; }

const char* esd_strerror(int s) {
    switch (s) {
        case ESD_OK: return "ok";
        case ESD_ERR_INVALID: return "invalid argument";
        case ESD_ERR_CUDA: return "CUDA error";
        case ESD_ERR_NOMEM: return "out of memory";
        case ESD_ERR_STATE: return "call out of order";
        case ESD_ERR_UNSUPPORTED: return "unsupported configuration";
        case ESD_ERR_CAPACITY: return "buffer too small";
        default: return "unknown status";
    }
}

const char* esd_last_error(const esd_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int esd_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

void esd_config_default(esd_config* cfg) {
    memset(cfg, 0, sizeof *cfg);
    cfg->struct_size = sizeof *cfg;
    cfg->detectors = ESD_DET_CONTENT;
    cfg->content_threshold = 27.0;
    cfg->content_weights[0] = cfg->content_weights[1] = cfg->content_weights[2] = 1.0;
    cfg->content_weight_div = 3.0;
    cfg->content_min_scene_len = 15;
    cfg->content_filter_mode = ESD_FILTER_MERGE;
    cfg->adaptive_threshold = 3.0;
    cfg->adaptive_min_content_val = 15.0;
    cfg->adaptive_weights[0] = cfg->adaptive_weights[1] = cfg->adaptive_weights[2] = 1.0;
    cfg->adaptive_weight_div = 3.0;
    cfg->adaptive_window_width = 2;
    cfg->adaptive_min_scene_len = 15;
    cfg->hist_threshold = 0.05;
    cfg->hist_bins = 256;
    cfg->hist_min_scene_len = 15;
    cfg->thresh_threshold = 12.0;
    cfg->thresh_min_scene_len = 15;
    cfg->thresh_method = ESD_THRESH_FLOOR;
    cfg->hash_threshold = 0.395;
    cfg->hash_size = 16;
    cfg->hash_lowpass = 2;
    cfg->hash_min_scene_len = 15;
}

int esd_create(esd_ctx** out, const esd_config* cfg, int device) {
    if (!out || !cfg) return fail(nullptr, ESD_ERR_INVALID, "esd_create: null argument");
    *out = nullptr;
    if (cfg->struct_size != sizeof(esd_config))
        return fail(nullptr, ESD_ERR_INVALID, "esd_create: struct_size %u != %zu (ABI mismatch)", cfg->struct_size,
                    sizeof(esd_config));
    if (!(cfg->detectors & 31) || (cfg->detectors & ~31)) return fail(nullptr, ESD_ERR_INVALID, "esd_create: bad detector mask");
    if (cfg->detectors & ESD_DET_HASH) {
        const int64_t S = (int64_t)cfg->hash_size * cfg->hash_lowpass;
        if (cfg->hash_size < 1 || cfg->hash_lowpass < 1) return fail(nullptr, ESD_ERR_INVALID, "hash size and lowpass must be >= 1");
        if (S % 2) return fail(nullptr, ESD_ERR_UNSUPPORTED, "Odd-size DCT's are not implemented (size * lowpass = %lld)", (long long)S);
        if (S > 64 || cfg->hash_size * cfg->hash_size > 1024)
            return fail(nullptr, ESD_ERR_UNSUPPORTED, "hash: size * lowpass must be <= 64 and size * size <= 1024");
    }
    if (cfg->src_width < 1 || cfg->src_height < 1) return fail(nullptr, ESD_ERR_INVALID, "esd_create: bad frame size");
    if ((cfg->detectors & ESD_DET_HIST) && (cfg->hist_bins < 1 || cfg->hist_bins > 256))
        return fail(nullptr, ESD_ERR_UNSUPPORTED, "hist_bins must be in 1..256");
    if ((cfg->detectors & ESD_DET_ADAPTIVE) && cfg->adaptive_window_width < 1)
        return fail(nullptr, ESD_ERR_INVALID, "window_width must be at least 1.");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return fail(nullptr, ESD_ERR_CUDA, "no CUDA device available (libesd has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(nullptr, ESD_ERR_INVALID, "esd_create: device %d out of range", device);

    esd_ctx* c = new esd_ctx();
    c->cfg = *cfg;
    c->device = device;
    auto bail = [&](int rc) {
        g_create_error = c->err;
        esd_destroy(c);
        return rc;
    };
#define CUB(call)                                                                                      \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            fail(c, ESD_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));                     \
            return bail(ESD_ERR_CUDA);                                                                 \
        }                                                                                              \
    } while (0)
    CUB(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUB(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 9) {
        fail(c, ESD_ERR_UNSUPPORTED, "device sm_%d%d lacks cp.async.bulk; libesd targets sm_100a", prop.major, prop.minor);
        return bail(ESD_ERR_UNSUPPORTED);
    }
    c->num_sms = prop.multiProcessorCount;

    // ---- geometry: SceneManager auto-downscale target
    const int W = cfg->src_width, H = cfg->src_height;
    int dw = cfg->dst_width, dh = cfg->dst_height;
    if (dw <= 0 || dh <= 0) {
        dw = W; dh = H;
        if (W >= 256) {
            if (cfg->downscale_mode == ESD_DOWNSCALE_INT) {
                const int f = W / 256;
                if (f > 1) { dw = (int)std::max(1L, py_round((double)W / f)); dh = (int)std::max(1L, py_round((double)H / f)); }
            } else {
                const double f = W / 256.0;
                if (f > 1.0) { dw = (int)std::max(1L, py_round(W / f)); dh = (int)std::max(1L, py_round(H / f)); }
            }
        }
    }
    c->dst_w = dw; c->dst_h = dh;
    c->resize = !(dw == W && dh == H);
    c->i420 = cfg->src_format == ESD_FMT_I420;
    c->nv12 = cfg->src_format == ESD_FMT_NV12 || c->i420;
    if (cfg->src_format != ESD_FMT_BGR24 && cfg->src_format != ESD_FMT_NV12 && cfg->src_format != ESD_FMT_I420) {
        fail(c, ESD_ERR_INVALID, "esd_create: unknown src_format %d", cfg->src_format);
        return bail(ESD_ERR_INVALID);
    }
    if (c->nv12 && (!c->resize || (W & 1) || (H & 1) || H > 32766 || W > 8190)) {
        fail(c, ESD_ERR_UNSUPPORTED, "NV12 / I420 input needs even dimensions and a downscaling context (frames %dx%d -> %dx%d)", W, H, dw, dh);
        return bail(ESD_ERR_UNSUPPORTED);
    }
    c->row_bytes = c->nv12 ? W : W * 3;
    c->need_hash = (cfg->detectors & ESD_DET_HASH) != 0;
    // the hash detector's gray plane rides on the content pass (a hash-only context also produces the HSV sums)
    c->need_content = (cfg->detectors & (ESD_DET_CONTENT | ESD_DET_ADAPTIVE | ESD_DET_THRESHOLD | ESD_DET_HASH)) != 0;
    c->need_hist = (cfg->detectors & ESD_DET_HIST) != 0;
    c->need_edges = ((cfg->detectors & ESD_DET_CONTENT) && cfg->content_weights[3] > 0.0) ||
                    ((cfg->detectors & ESD_DET_ADAPTIVE) && cfg->adaptive_weights[3] > 0.0);
    c->extras = c->need_edges || c->need_hash || (cfg->detectors & ESD_DET_THRESHOLD) != 0;
    if (c->need_edges) {
        // _estimated_kernel_size: 4 + round(sqrt(w * h) / 192), made odd
        int ks = cfg->edge_kernel_size;
        if (ks <= 0) {
            ks = 4 + (int)py_round(sqrt((double)dw * (double)dh) / 192.0);
            if (ks % 2 == 0) ks += 1;
        }
        if (ks < 3 || ks % 2 == 0) {
            fail(c, ESD_ERR_INVALID, "kernel_size must be odd integer >= 3");
            return bail(ESD_ERR_INVALID);
        }
        if (ks > 63) {
            fail(c, ESD_ERR_UNSUPPORTED, "kernel_size %d too large for the bit-packed dilation (max 63)", ks);
            return bail(ESD_ERR_UNSUPPORTED);
        }
        c->edge_ksize = ks;
        c->edge_words = dh * ((dw + 31) / 32);  // row-padded bitmaps
        const size_t esmem = 2 * (size_t)((dw * dh + 31) & ~31) + 8 * (size_t)c->edge_words;
        if (esmem + 2048 > prop.sharedMemPerBlockOptin) {
            fail(c, ESD_ERR_UNSUPPORTED, "delta_edges needs the detector-resolution frame (%dx%d) in shared memory; at most ~%zu pixels",
                 dw, dh, (size_t)(prop.sharedMemPerBlockOptin - 2048) / 2);
            return bail(ESD_ERR_UNSUPPORTED);
        }
        CUB(cudaFuncSetAttribute(edges_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)esmem));
        CUB(esdguard::gmalloc(&c->d_edge_prev, sizeof(uint32_t) * c->edge_words));
    }
    if (c->need_hash) {
        HashParams& hp = c->hparams;
        hp.w = dw; hp.h = dh;
        hp.hs = cfg->hash_size;
        hp.S = cfg->hash_size * cfg->hash_lowpass;
        if (hp.S > dw || hp.S > dh) {
            fail(c, ESD_ERR_UNSUPPORTED, "hash: the %dx%d DCT thumbnail is larger than the detector-resolution frame (%dx%d)", hp.S, hp.S, dw, dh);
            return bail(ESD_ERR_UNSUPPORTED);
        }
        c->hash_words = (hp.hs * hp.hs + 31) / 32;
        hp.words = c->hash_words;
        const double sx = (double)dw / hp.S, sy = (double)dh / hp.S;
        hp.isx = (int)sx; hp.isy = (int)sy;
        hp.fast = (fabs(sx - hp.isx) < 2.220446049250313e-16 && fabs(sy - hp.isy) < 2.220446049250313e-16) ? 1 : 0;
        hp.fast_scale = (float)(1.0 / (hp.isx * hp.isy));
        std::vector<int> xb, xs, yb, ys;
        std::vector<float> xw, yw;
        area_axis_table(dw, hp.S, xb, xs, xw);
        area_axis_table(dh, hp.S, yb, ys, yw);
        for (int d = 0; d < hp.S; ++d)  // hash_kernel walks the taps of a destination column as consecutive pixels
            for (int k = xb[d] + 1; k < xb[d + 1]; ++k)
                if (xs[k] != xs[k - 1] + 1) {
                    fail(c, ESD_ERR_UNSUPPORTED, "hash: non-contiguous area-resize taps for %d -> %d", dw, hp.S);
                    return bail(ESD_ERR_UNSUPPORTED);
                }
        // horizontal pass as an exact integer sum: the fast path, or one power-of-two weight everywhere (e.g. 256 -> 32)
        bool uniform = true;
        for (float v : xw) uniform = uniform && v == xw[0];
        int expo = 0;
        const bool pow2 = uniform && frexpf(xw[0], &expo) == 0.5f;
        hp.xsum_int = (hp.fast || pow2) ? 1 : 0;
        hp.xalpha = xw[0];
        // aligned-words variant: uniform tap count, a multiple of 4 bytes, every tap run on a 4-byte boundary of the plane
        hp.xwords = 0;
        if (hp.xsum_int && dw % 4 == 0 && ((int64_t)dw * dh) % 4 == 0) {
            const int cnt = xb[1] - xb[0];
            bool ok = cnt % 4 == 0 && cnt / 4 >= 1 && cnt / 4 <= 4;
            for (int d = 0; d < hp.S && ok; ++d) ok = (xb[d + 1] - xb[d] == cnt) && (xs[xb[d]] % 4 == 0);
            if (ok) hp.xwords = cnt / 4;
        }
        hp.n_xent = (int)xs.size();
        hp.n_yent = (int)ys.size();
        // band buffer of the area resize: at least the source rows of one destination row
        int need = 1;
        for (int d = 0; d < hp.S; ++d) need = std::max(need, ys[yb[d + 1] - 1] - ys[yb[d]] + 1);
        hp.band_rows = std::min(dh, std::max(need, (12 * 1024) / (hp.S * 4)));  // 12 KB: keeps eight CTAs per SM resident
        std::vector<int> bands;  // {dy0, dy1, r0, nrows}
        for (int dy0 = 0; dy0 < hp.S;) {
            const int r0 = ys[yb[dy0]];
            int dy1 = dy0 + 1;
            while (dy1 < hp.S && ys[yb[dy1 + 1] - 1] - r0 + 1 <= hp.band_rows) ++dy1;
            bands.insert(bands.end(), {dy0, dy1, r0, ys[yb[dy1] - 1] - r0 + 1});
            dy0 = dy1;
        }
        hp.n_bands = (int)bands.size() / 4;
        std::vector<int> itab;
        itab.insert(itab.end(), xb.begin(), xb.end());
        for (int d = 0; d < hp.S; ++d) itab.push_back(xs[xb[d]]);
        itab.insert(itab.end(), yb.begin(), yb.end());
        itab.insert(itab.end(), ys.begin(), ys.end());
        while (itab.size() % 4) itab.push_back(0);
        itab.insert(itab.end(), bands.begin(), bands.end());
        std::vector<float> wtab(xw);
        wtab.insert(wtab.end(), yw.begin(), yw.end());
        CUB(esdguard::gmalloc(&c->d_hash_itab, sizeof(int) * itab.size()));
        CUB(cudaMemcpy(c->d_hash_itab, itab.data(), sizeof(int) * itab.size(), cudaMemcpyHostToDevice));
        CUB(esdguard::gmalloc(&c->d_hash_wtab, sizeof(float) * wtab.size()));
        CUB(cudaMemcpy(c->d_hash_wtab, wtab.data(), sizeof(float) * wtab.size(), cudaMemcpyHostToDevice));
        hp.itab = c->d_hash_itab;
        hp.wtab = c->d_hash_wtab;
        // orthonormal DCT-II rows 0..hs-1 of size S
        std::vector<double> Cm((size_t)hp.hs * hp.S);
        const double pi = 3.14159265358979323846;
        for (int u = 0; u < hp.hs; ++u)
            for (int i = 0; i < hp.S; ++i)
                Cm[(size_t)u * hp.S + i] = u == 0 ? sqrt(1.0 / hp.S) : sqrt(2.0 / hp.S) * cos(pi * (2 * i + 1) * u / (2.0 * hp.S));
        CUB(esdguard::gmalloc(&c->d_hash_C, sizeof(double) * Cm.size()));
        CUB(cudaMemcpy(c->d_hash_C, Cm.data(), sizeof(double) * Cm.size(), cudaMemcpyHostToDevice));
        hp.C = c->d_hash_C;
        {   // shared-memory carve-up of hash_kernel
            const size_t region = std::max<size_t>(sizeof(float) * hp.band_rows * hp.S,
                                                   sizeof(double) * hp.hs * hp.S + sizeof(uint32_t) * hp.hs * hp.hs);
            c->hash_smem = sizeof(double) * hp.hs * hp.S + sizeof(float) * hp.S * hp.S + ((region + 7) & ~(size_t)7) +
                           sizeof(float) * (hp.n_xent + hp.n_yent) + sizeof(int) * (3 * hp.S + 2 + hp.n_yent) + 16;
        }
        if (c->hash_smem + 2048 > prop.sharedMemPerBlockOptin) {
            fail(c, ESD_ERR_UNSUPPORTED, "hash: frame too large for the shared-memory tables of the area resize (%zu bytes)", c->hash_smem);
            return bail(ESD_ERR_UNSUPPORTED);
        }
        c->hash_fn = hp.xwords == 1 ? hash_kernel<1> : hp.xwords == 2 ? hash_kernel<2> : hp.xwords == 3 ? hash_kernel<3>
                   : hp.xwords == 4 ? hash_kernel<4> : hash_kernel<0>;
        CUB(cudaFuncSetAttribute(c->hash_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->hash_smem));
    }
    if (dw > kConsumers * (c->resize ? 4 : 16)) {
        fail(c, ESD_ERR_UNSUPPORTED, "destination width %d too large (max %d)", dw, kConsumers * (c->resize ? 4 : 16));
        return bail(ESD_ERR_UNSUPPORTED);
    }
    int pxt = (dw + kConsumers - 1) / kConsumers;
    c->pxt = pxt <= 1 ? 1 : pxt <= 2 ? 2 : pxt <= 4 ? 4 : pxt <= 8 ? 8 : 16;

    std::vector<int> xo0, xo1, xa0, xa1, yo0, yo1, yb0, yb1;
    if (c->resize) {
        axis_tables(W, dw, xo0, xo1, xa0, xa1);
        axis_tables(H, dh, yo0, yo1, yb0, yb1);
    } else {
        yo0.resize(dh); yo1.resize(dh); yb0.assign(dh, 2048); yb1.assign(dh, 0);
        for (int y = 0; y < dh; ++y) yo0[y] = yo1[y] = y;
    }
    {  // touched rows and compact indices
        std::vector<char> used(H, 0);
        for (int y = 0; y < dh; ++y) { used[yo0[y]] = 1; if (c->resize) used[yo1[y]] = 1; }
        std::vector<int> cidx(H, -1);
        for (int r = 0; r < H; ++r)
            if (used[r]) { cidx[r] = (int)c->touched.size(); c->touched.push_back(r); }
        c->n_touched_y = (int)c->touched.size();
        std::vector<int> uvidx(H / 2 + 1, -1);
        if (c->nv12) {  // UV rows behind the Y rows, as virtual rows H + r of a contiguous NV12 frame
            std::vector<char> uused(H / 2, 0);
            for (int y = 0; y < dh; ++y) { uused[yo0[y] >> 1] = 1; uused[yo1[y] >> 1] = 1; }
            for (int r = 0; r < H / 2; ++r)
                if (uused[r]) { uvidx[r] = (int)c->touched.size(); c->touched.push_back(H + r); }
        }
        std::vector<YRow> yr(dh);
        for (int y = 0; y < dh; ++y) {
            yr[y].row0 = yo0[y]; yr[y].row1 = yo1[y];
            yr[y].crow0 = cidx[yo0[y]]; yr[y].crow1 = c->resize ? cidx[yo1[y]] : cidx[yo0[y]];
            yr[y].b0s = (uint32_t)yb0[y] << 16; yr[y].b1s = (uint32_t)yb1[y] << 16;
            yr[y].uvrows = yr[y].cuvrows = 0;
            if (c->nv12) {
                yr[y].uvrows = (uint32_t)(yo0[y] >> 1) | ((uint32_t)(yo1[y] >> 1) << 16);
                yr[y].cuvrows = (uint32_t)uvidx[yo0[y] >> 1] | ((uint32_t)uvidx[yo1[y] >> 1] << 16);
            }
        }
        if (c->i420) {
            std::vector<int> rows;
            for (size_t i = (size_t)c->n_touched_y; i < c->touched.size(); ++i) rows.push_back(c->touched[i] - H);
            CUB(esdguard::gmalloc(&c->d_uv_src_rows, sizeof(int) * std::max<size_t>(1, rows.size())));
            CUB(cudaMemcpy(c->d_uv_src_rows, rows.data(), sizeof(int) * rows.size(), cudaMemcpyHostToDevice));
        }
        CUB(esdguard::gmalloc(&c->d_yrows, sizeof(YRow) * dh));
        CUB(cudaMemcpy(c->d_yrows, yr.data(), sizeof(YRow) * dh, cudaMemcpyHostToDevice));
    }
    {
        std::vector<uint2> xt(dw);
        for (int x = 0; x < dw; ++x) {
            if (c->resize && c->nv12) {  // luma byte offset | chroma pair byte offset << 13 | tap 1 in the next pair << 26
                const uint32_t x0 = (uint32_t)xo0[x];
                xt[x].x = x0 | ((x0 & ~1u) << 13) | ((x0 & 1u) << 26);
                xt[x].y = (uint32_t)xa0[x] | ((uint32_t)xa1[x] << 16);
            } else if (c->resize) { xt[x].x = (uint32_t)(3 * xo0[x]); xt[x].y = (uint32_t)xa0[x] | ((uint32_t)xa1[x] << 16); }
            else { xt[x].x = (uint32_t)(3 * x); xt[x].y = 2048u; }
        }
        {  // sector-granular touched bytes per row (SURVEY.md section 8d accounting)
            std::vector<char> sec((c->row_bytes + 31) / 32, 0);
            for (int x = 0; x < dw && !c->nv12; ++x) {  // (NV12's x table is packed differently: handled below)
                const int b0 = (int)xt[x].x, b1 = std::min(c->row_bytes, b0 + (c->resize ? 6 : 3)) - 1;
                for (int q = b0 / 32; q <= b1 / 32; ++q) sec[q] = 1;
            }
            if (c->nv12) {  // Y rows: bytes x0, x0 + 1; UV rows: the chroma pairs of those two columns
                std::fill(sec.begin(), sec.end(), 0);
                std::vector<char> secuv(sec.size(), 0);
                for (int x = 0; x < dw; ++x) {
                    const int x0 = xo0[x], x1 = std::min(x0 + 1, W - 1);
                    sec[x0 / 32] = sec[x1 / 32] = 1;
                    secuv[(x0 & ~1) / 32] = secuv[((x0 & ~1) + 1) / 32] = 1;
                    secuv[(x1 & ~1) / 32] = secuv[((x1 & ~1) + 1) / 32] = 1;
                }
                int ny = 0, nuv = 0;
                for (char v : sec) ny += v;
                for (char v : secuv) nuv += v;
                c->alg_row_bytes = std::min(c->row_bytes, ny * 32);
                c->alg_frame_bytes = (int64_t)c->n_touched_y * c->alg_row_bytes +
                                     (int64_t)(c->touched.size() - c->n_touched_y) * std::min(c->row_bytes, nuv * 32);
            } else {
                int nsec = 0;
                for (char v : sec) nsec += v;
                c->alg_row_bytes = std::min(c->row_bytes, nsec * 32);
                c->alg_frame_bytes = (int64_t)c->touched.size() * c->alg_row_bytes;
            }
        }
        CUB(esdguard::gmalloc(&c->d_xtab, sizeof(uint2) * dw));
        CUB(cudaMemcpy(c->d_xtab, xt.data(), sizeof(uint2) * dw, cudaMemcpyHostToDevice));
        if (c->resize) {
            // tap-compact layout.  BGR24: column d of a gathered row holds [tap0 BGR, tap1 BGR] at byte 6 d.  NV12: a gathered
            // Y row holds [Y(x0), Y(x0+1)] at byte 2 d, a gathered UV row [U V of tap 0, U V of tap 1] at byte 4 d (one pitch).
            c->tap_row_bytes = ((c->nv12 ? 4 : 6) * dw + 15) & ~15;
            c->tap_src_off.resize(dw);
            std::vector<uint2> xt2(dw);
            for (int x = 0; x < dw; ++x) {
                c->tap_src_off[x] = (c->nv12 ? 1 : 3) * xo0[x];
                xt2[x].x = c->nv12 ? ((uint32_t)(2 * x) | ((uint32_t)(4 * x) << 13) | (1u << 26)) : (uint32_t)(6 * x);
                xt2[x].y = xt[x].y;
            }
            CUB(esdguard::gmalloc(&c->d_xtab_taps, sizeof(uint2) * dw));
            CUB(cudaMemcpy(c->d_xtab_taps, xt2.data(), sizeof(uint2) * dw, cudaMemcpyHostToDevice));
            // bank-aware lane -> column mapping (BGR24: three words per row and tap pair; NV12: two luma words -- its chroma
            // words follow the same lattice at half the density)
            std::vector<uint32_t> off(dw), off2(dw);
            for (int x = 0; x < dw; ++x) { off[x] = xt[x].x & (c->nv12 ? 0x1fffu : 0xffffffffu); off2[x] = xt2[x].x & (c->nv12 ? 0x1fffu : 0xffffffffu); }
            const int forced = cfg->reserved1;  // tuning / A-B switch: 1, 2, 4, 8 force a stride; 0 = choose
            const bool ok = forced == 1 || forced == 2 || forced == 4 || forced == 8;
            c->lane_stride = ok ? forced : choose_lane_stride(off, dw, c->nv12 ? 2 : 3);
            c->lane_stride_taps = ok ? forced : choose_lane_stride(off2, dw, c->nv12 ? 2 : 3);
        }
    }
    {  // OpenCV RGB2HSV_b tables (A.3)
        int sdiv[256], hdiv[256];
        sdiv[0] = hdiv[0] = 0;
        for (int i = 1; i < 256; ++i) {
            sdiv[i] = (int)lrint((255 << 12) / (1. * i));
            hdiv[i] = (int)lrint((180 << 12) / (6. * i));
        }
        CUB(esdguard::gmalloc(&c->d_sdiv, sizeof sdiv));
        CUB(esdguard::gmalloc(&c->d_hdiv, sizeof hdiv));
        CUB(cudaMemcpy(c->d_sdiv, sdiv, sizeof sdiv, cudaMemcpyHostToDevice));
        CUB(cudaMemcpy(c->d_hdiv, hdiv, sizeof hdiv, cudaMemcpyHostToDevice));
    }

    // ---- kernel shape
    c->rowbuf = ((c->row_bytes + 15 + 15) & ~15) + 16;
    // destination rows per pipeline stage: 2 amortises the per-stage barrier/metadata work over two pixels per thread
    // full-resolution scoring (no resize, >= 4 pixels per thread per row) is compute/latency-bound: small groups,
    // single-row stages and a deeper ring measured best there (scripts/noresize_probe.py)
    const bool wide_noresize = !c->resize && c->pxt >= 4;
    int RS = cfg->rows_per_stage > 0 ? cfg->rows_per_stage : (wide_noresize ? 1 : 4);
    RS = std::max(1, std::min(RS, kMaxRowsPerStage));
    const int rows_per_dst = c->resize ? (c->nv12 ? 4 : 2) : 1;  // staged source rows per destination row
    while (RS > 1 && (size_t)RS * rows_per_dst * c->rowbuf > 48 * 1024) --RS;
    c->rows_per_stage = RS;
    c->stage_bytes = RS * rows_per_dst * c->rowbuf;
    // rows per group: 16 measured best at 1080p->256x144 (profiles/r01_sweep.md); the previous-frame HSV of
    // a group lives in shared memory (R * pxt KB), keep it <= 32 KB unless the caller insists
    int R = cfg->rows_per_group > 0 ? cfg->rows_per_group : wide_noresize ? std::max(1, std::min(4, 32 / c->pxt)) : std::max(1, std::min(16, 32 / c->pxt));
    if (cfg->rows_per_group <= 0 && !wide_noresize && R == 16 && c->need_content) {
        // Narrow source rows (720p BGR24, 1080p NV12) leave room for a third CTA per SM if the previous-frame HSV of a group
        // takes 8 KB instead of 16: the consumer warps are latency-bound there (issue slots 60 % busy with ~4 warps per
        // scheduler, profiles/r02_fused_ncu.md), so 27 resident warps beat 18 -- 720p 0.84 -> 0.90 of the HBM peak
        // (profiles/r02_kernel_ab_paired_rows_shapes.log).  1080p / 4K BGR24 rows fit two CTAs either way and keep 16.
        const int S0 = cfg->pipeline_stages > 0 ? std::min(cfg->pipeline_stages, kMaxStages) : (c->rows_per_stage >= 4 ? 2 : c->rows_per_stage > 1 ? 3 : 4);
        auto ctas_for = [&](int r) {
            const size_t smem = fused_smem_bytes(c, r, S0) + 1024;  // + the per-CTA reservation
            return (int)std::min<size_t>(prop.sharedMemPerMultiprocessor / smem, (size_t)(prop.maxThreadsPerMultiProcessor / kThreads));
        };
        if (ctas_for(8) > ctas_for(16)) R = 8;
    }
    R = std::min(R, dh);
    R = std::min(R, 255);
    while (R > 1 && ((int64_t)R * c->pxt > 256 || (int64_t)R * dw > 65535)) --R;
    c->rows_per_group = R;
    c->n_groups = (dh + R - 1) / R;
    int stages = cfg->pipeline_stages > 0 ? std::min(cfg->pipeline_stages, kMaxStages) : (wide_noresize ? 4 : c->rows_per_stage >= 4 ? 2 : c->rows_per_stage > 1 ? 3 : 4);
    const size_t smem_limit = prop.sharedMemPerBlockOptin;
    while (stages > 2 && fused_smem_bytes(c, R, stages) > smem_limit) --stages;
    c->stages = stages;
    c->smem_bytes = fused_smem_bytes(c, R, stages);
    if (c->smem_bytes > smem_limit) {
        fail(c, ESD_ERR_UNSUPPORTED, "frame rows too wide for shared memory staging (%zu > %zu bytes)", c->smem_bytes, smem_limit);
        return bail(ESD_ERR_UNSUPPORTED);
    }
    int occ = 0;
    CUB(ESD_DISPATCH(occupancy_rp, c->resize, c->pxt, c->need_content, c->need_hist, c->nv12, c->extras, c->smem_bytes, &occ));
    if (occ < 1) {
        fail(c, ESD_ERR_UNSUPPORTED, "fused kernel does not fit on an SM (smem %zu)", c->smem_bytes);
        return bail(ESD_ERR_UNSUPPORTED);
    }
    c->ctas_per_sm = cfg->ctas_per_sm > 0 ? std::min(cfg->ctas_per_sm, occ) : occ;

    // ---- scoring / decision parameters
    auto weights = [](const double* w, double div, ScoreWeights* o) {
        for (int i = 0; i < 4; ++i) o->w[i] = w[i];
        if (div > 0) o->div = div;
        else o->div = ((fabs(w[0]) + fabs(w[1])) + fabs(w[2])) + fabs(w[3]);
    };
    weights(cfg->content_weights, cfg->content_weight_div, &c->wc);
    weights(cfg->adaptive_weights, cfg->adaptive_weight_div, &c->wa);
    c->max_cuts = cfg->max_cuts > 0 ? cfg->max_cuts : 65536;
    DecisionParams& P = c->dparams;
    P.detectors = cfg->detectors;
    P.content_min_scene_len = cfg->content_min_scene_len;
    P.content_filter_mode = cfg->content_filter_mode;
    P.adaptive_w = cfg->adaptive_window_width;
    P.adaptive_min_scene_len = cfg->adaptive_min_scene_len;
    P.hist_min_scene_len = cfg->hist_min_scene_len;
    P.content_threshold = cfg->content_threshold;
    P.adaptive_threshold = cfg->adaptive_threshold;
    P.adaptive_min_content_val = cfg->adaptive_min_content_val;
    P.hist_threshold = std::max(0.0, std::min(1.0, 1.0 - cfg->hist_threshold));
    P.max_cuts = c->max_cuts;
    P.thresh_threshold = (double)(long long)cfg->thresh_threshold;  // self.threshold = int(threshold)
    P.thresh_fade_bias = cfg->thresh_fade_bias;
    P.thresh_min_scene_len = cfg->thresh_min_scene_len;
    P.thresh_method = cfg->thresh_method;
    P.hash_threshold = cfg->hash_threshold;
    P.hash_min_scene_len = cfg->hash_min_scene_len;
    P.cuts_stride = c->max_cuts;

    CUB(esdguard::gmalloc(&c->d_state, sizeof(DecisionState)));
    CUB(esdguard::gmalloc(&c->d_cuts, sizeof(long long) * 5 * c->max_cuts));
    if (c->need_content) {
        CUB(esdguard::gmalloc(&c->d_prev[0], sizeof(uint32_t) * dw * dh));
        CUB(esdguard::gmalloc(&c->d_prev[1], sizeof(uint32_t) * dw * dh));
    }
    CUB(cudaEventCreateWithFlags(&c->order_event, cudaEventDisableTiming));
    CUB(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
    CUB(cudaEventCreateWithFlags(&c->ev_fused, cudaEventDisableTiming));
    CUB(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    CUB(cudaEventCreateWithFlags(&c->ev_fin[0], cudaEventDisableTiming));
    CUB(cudaEventCreateWithFlags(&c->ev_fin[1], cudaEventDisableTiming));
    if (ensure_capacity(c, std::max<int64_t>(4096, cfg->initial_capacity)) != ESD_OK) return bail(ESD_ERR_CUDA);
    if (reset_video_state(c) != ESD_OK) return bail(ESD_ERR_CUDA);
#undef CUB
    *out = c;
    return ESD_OK;
}

void esd_destroy(esd_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    sync_all(c);  // this context's streams only: destroying one context must not stall the others on the device
    if (c->pf_stream) cudaStreamSynchronize(c->pf_stream);
    if (c->dec_stream) cudaStreamSynchronize(c->dec_stream);
    esd_ingest_close(c);
    free_plans(c);
    for (auto& ev : c->timing_events) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    esdguard::gfree(c->d_xtab_taps);
    esdguard::gfree(c->d_uv_src_rows); esdguard::gfree(c->d_uvpack);
    esdguard::gfree(c->d_yrows); esdguard::gfree(c->d_xtab); esdguard::gfree(c->d_sdiv); esdguard::gfree(c->d_hdiv);
    esdguard::gfree(c->d_prev[0]); esdguard::gfree(c->d_prev[1]); esdguard::gfree(c->d_state); esdguard::gfree(c->d_cuts);
    esdguard::gfree(c->d_slab);
    for (int b = 0; b < 2; ++b) { esdguard::gfree(c->d_part[b]); esdguard::gfree(c->d_hist_part[b]); esdguard::gfree(c->d_vplane[b]); if (c->ev_fin[b]) cudaEventDestroy(c->ev_fin[b]); }
    esdguard::gfree(c->d_edge_bits); esdguard::gfree(c->d_edge_prev);
    c->kg.destroy();
    esdguard::gfree(c->dec_d_scores); esdguard::gfree(c->dec_d_ratio); esdguard::gfree(c->dec_d_state);
    if (c->dec_h) cudaFreeHost(c->dec_h);
    if (c->dec_stream) cudaStreamDestroy(c->dec_stream);
    if (c->pf_h) cudaFreeHost(c->pf_h);
    if (c->pf_mailbox) cudaFreeHost(c->pf_mailbox);
    if (c->pf_stream) cudaStreamDestroy(c->pf_stream);
    esdguard::gfree(c->d_gplane[0]); esdguard::gfree(c->d_gplane[1]); esdguard::gfree(c->d_hash_small);
    esdguard::gfree(c->d_hash_itab); esdguard::gfree(c->d_hash_wtab); esdguard::gfree(c->d_hash_C);
    if (c->order_event) cudaEventDestroy(c->order_event);
    if (c->ev_fused) cudaEventDestroy(c->ev_fused);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
    cudaGetLastError();
    delete c;
}

int esd_reset(esd_ctx* c) {
    if (!c) return ESD_ERR_INVALID;
    CU(c, cudaSetDevice(c->device));
    { int rc = sync_all(c); if (rc) return rc; }
    if (c->pf_stream) CU(c, cudaStreamSynchronize(c->pf_stream));
    if (c->compute_stream) CU(c, cudaStreamSynchronize(c->compute_stream));
    return reset_video_state(c);
}

int esd_get_geometry(const esd_ctx* c, esd_geometry* g) {
    if (!c || !g) return ESD_ERR_INVALID;
    g->dst_width = c->dst_w;
    g->dst_height = c->dst_h;
    g->n_touched_rows = (int32_t)c->touched.size();
    g->row_bytes = c->row_bytes;
    g->alg_bytes_per_frame = c->alg_frame_bytes;
    g->compact_frame_bytes = (int64_t)c->touched.size() * c->row_bytes;
    g->lane_stride = c->lane_stride;
    g->lane_stride_taps = c->lane_stride_taps;
    return ESD_OK;
}

int esd_get_touched_rows(const esd_ctx* c, int32_t* rows, int32_t cap) {
    if (!c || !rows) return ESD_ERR_INVALID;
    if (cap < (int32_t)c->touched.size()) return ESD_ERR_CAPACITY;
    memcpy(rows, c->touched.data(), sizeof(int32_t) * c->touched.size());
    return ESD_OK;
}

int esd_push_frames(esd_ctx* c, const uint8_t* d_bgr, int64_t n, int64_t frame_stride, int64_t pitch,
                    int64_t first_frame_num, void* stream) {
    if (!c) return ESD_ERR_INVALID;
    if (c->i420) {  // contiguous I420 frames: Y plane (pitch), then the U and the V plane (pitch / 2 each, src_height / 2 rows)
        if (pitch & 1) return fail(c, ESD_ERR_INVALID, "push: I420 frames need an even pitch (chroma rows are pitch / 2 apart)");
        const uint8_t* d_u = d_bgr ? d_bgr + (int64_t)c->cfg.src_height * pitch : nullptr;
        return esd_push_i420(c, d_bgr, d_u, d_u ? d_u + (int64_t)(c->cfg.src_height / 2) * (pitch / 2) : nullptr, n, frame_stride, pitch,
                             pitch / 2, first_frame_num, stream);
    }
    if (c->nv12)  // contiguous NV12 frames: the UV plane starts at row src_height of the same pitch
        return esd_push_nv12(c, d_bgr, d_bgr ? d_bgr + (int64_t)c->cfg.src_height * pitch : nullptr, n, frame_stride, pitch,
                             first_frame_num, stream);
    if (pitch < c->row_bytes) return fail(c, ESD_ERR_INVALID, "push: pitch %lld < row bytes %d", (long long)pitch, c->row_bytes);
    if (n > 1 && frame_stride < pitch * (int64_t)(c->cfg.src_height - 1) + c->row_bytes)
        return fail(c, ESD_ERR_INVALID, "push: frame stride %lld smaller than a frame", (long long)frame_stride);
    if (n > 0 && d_bgr) {
        const size_t span = (size_t)(n - 1) * (size_t)frame_stride + (size_t)(c->cfg.src_height - 1) * (size_t)pitch + (size_t)c->row_bytes;
        int rc = validate_device_span(c, d_bgr, span, "push_frames");
        if (rc) return rc;
    }
    return push_common(c, d_bgr, n, frame_stride, pitch, LAYOUT_FULL, first_frame_num, (cudaStream_t)stream);
}

int esd_push_nv12(esd_ctx* c, const uint8_t* d_y, const uint8_t* d_uv, int64_t n, int64_t frame_stride, int64_t pitch,
                  int64_t first_frame_num, void* stream) {
    if (!c) return ESD_ERR_INVALID;
    if (!c->nv12 || c->i420) return fail(c, ESD_ERR_STATE, "push_nv12: the context was not created for NV12 frames (src_format)");
    if (!d_y || !d_uv) return fail(c, ESD_ERR_INVALID, "push_nv12: null plane");
    if (pitch < c->row_bytes) return fail(c, ESD_ERR_INVALID, "push_nv12: pitch %lld < row bytes %d", (long long)pitch, c->row_bytes);
    const int H = c->cfg.src_height;
    if (n > 0) {
        const size_t fspan = (size_t)std::max<int64_t>(0, n - 1) * (size_t)frame_stride;
        int rc = validate_device_span(c, d_y, fspan + (size_t)(H - 1) * (size_t)pitch + (size_t)c->row_bytes, "push_nv12 (Y plane)");
        if (rc) return rc;
        rc = validate_device_span(c, d_uv, fspan + (size_t)(H / 2 - 1) * (size_t)pitch + (size_t)c->row_bytes, "push_nv12 (UV plane)");
        if (rc) return rc;
    }
    return push_common(c, d_y, n, frame_stride, pitch, LAYOUT_FULL, first_frame_num, (cudaStream_t)stream, false, nullptr, d_uv);
}

// launches i420_interleave_kernel for n frames: chroma rows from (u, v) -> interleaved rows at dst
static int interleave_chroma(esd_ctx* c, const uint8_t* u, const uint8_t* v, int64_t src_fs, int64_t src_pitch, const int* src_rows,
                             uint8_t* dst, int64_t dst_fs, int64_t dst_pitch, int64_t n, cudaStream_t st) {
    const int n_uv = (int)c->touched.size() - c->n_touched_y;
    const int half_w = c->cfg.src_width / 2;
    for (int64_t f0 = 0; f0 < n; f0 += 65535) {  // gridDim.y limit
        const int64_t m = std::min<int64_t>(65535, n - f0);
        CU(c, klaunch(c->kg, st, i420_interleave_kernel, dim3((unsigned)n_uv, (unsigned)m), dim3(kInterleaveThreads), (size_t)(2 * half_w),
                      u + f0 * src_fs, v + f0 * src_fs, (long long)src_fs, (long long)src_pitch, src_rows, dst + f0 * dst_fs,
                      (long long)dst_fs, (long long)dst_pitch, half_w));
        c->launches++;
    }
    return ESD_OK;
}

int esd_push_i420(esd_ctx* c, const uint8_t* d_y, const uint8_t* d_u, const uint8_t* d_v, int64_t n, int64_t frame_stride,
                  int64_t pitch_y, int64_t pitch_uv, int64_t first_frame_num, void* stream) {
    if (!c) return ESD_ERR_INVALID;
    if (!c->i420) return fail(c, ESD_ERR_STATE, "push_i420: the context was not created for I420 frames (src_format)");
    if (!d_y || !d_u || !d_v) return fail(c, ESD_ERR_INVALID, "push_i420: null plane");
    if (n <= 0) return fail(c, ESD_ERR_INVALID, "push: null frames or n <= 0");
    if (pitch_y < c->row_bytes || pitch_uv < c->row_bytes / 2)
        return fail(c, ESD_ERR_INVALID, "push_i420: pitch %lld / %lld < row bytes %d / %d", (long long)pitch_y, (long long)pitch_uv,
                    c->row_bytes, c->row_bytes / 2);
    const int H = c->cfg.src_height;
    const size_t fspan = (size_t)(n - 1) * (size_t)frame_stride;
    int rc = validate_device_span(c, d_y, fspan + (size_t)(H - 1) * (size_t)pitch_y + (size_t)c->row_bytes, "push_i420 (Y plane)");
    if (rc) return rc;
    const size_t cspan = fspan + (size_t)(H / 2 - 1) * (size_t)pitch_uv + (size_t)(c->row_bytes / 2);
    if ((rc = validate_device_span(c, d_u, cspan, "push_i420 (U plane)"))) return rc;
    if ((rc = validate_device_span(c, d_v, cspan, "push_i420 (V plane)"))) return rc;
    if (c->poisoned) return fail(c, ESD_ERR_STATE, "push: an earlier push failed half-way; call esd_reset");
    CU(c, cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n_uv = (int64_t)c->touched.size() - c->n_touched_y;
    const int64_t uv_pitch = (c->row_bytes + 15) & ~15;
    const size_t need = (size_t)n * (size_t)n_uv * (size_t)uv_pitch;
    if (need > c->uvpack_bytes) {
        if ((rc = sync_all(c))) return rc;  // an earlier push may still read the old buffer
        esdguard::gfree(c->d_uvpack);
        c->d_uvpack = nullptr;
        c->uvpack_bytes = 0;
        CU(c, esdguard::gmalloc(&c->d_uvpack, need));
        c->uvpack_bytes = need;
    }
    // the repack writes the buffer the previous push's fused kernel reads: order this stream behind it first
    if ((rc = order_after_last(c, st))) return rc;
    if ((rc = interleave_chroma(c, d_u, d_v, frame_stride, pitch_uv, c->d_uv_src_rows, c->d_uvpack, n_uv * uv_pitch, uv_pitch, n, st))) return rc;
    const UvPacked uvp{n_uv * uv_pitch, uv_pitch};
    return push_common(c, d_y, n, frame_stride, pitch_y, LAYOUT_FULL, first_frame_num, st, false, nullptr, c->d_uvpack, &uvp);
}

int esd_push_rows(esd_ctx* c, const uint8_t* d_rows, int64_t n, int64_t first_frame_num, void* stream) {
    if (!c) return ESD_ERR_INVALID;
    if (n > 0 && d_rows) {
        int rc = validate_device_span(c, d_rows, (size_t)n * c->touched.size() * c->row_bytes, "push_rows");
        if (rc) return rc;
    }
    return push_common(c, d_rows, n, (int64_t)c->touched.size() * c->row_bytes, c->row_bytes, LAYOUT_ROWS, first_frame_num,
                       (cudaStream_t)stream);
}

static int ensure_pinned(esd_ctx* c, IngestSlot& s, size_t bytes) {
    if (s.h_pinned && s.h_pinned_bytes >= bytes) return ESD_OK;
    if (s.h_pinned) cudaFreeHost(s.h_pinned);
    s.h_pinned = nullptr;
    s.h_pinned_bytes = 0;
    CU(c, cudaHostAlloc(&s.h_pinned, bytes, cudaHostAllocDefault));
    s.h_pinned_bytes = bytes;
    return ESD_OK;
}

// ------------------------------------------------------------------------------------- ingest ring
int esd_ingest_open(esd_ctx* c, int32_t n_slots, int32_t frames_per_slot) {
    if (!c) return ESD_ERR_INVALID;
    if (!c->ring.empty()) return fail(c, ESD_ERR_STATE, "ingest ring already open");
    if (n_slots < 2 || frames_per_slot < 1) return fail(c, ESD_ERR_INVALID, "ingest: need >= 2 slots and >= 1 frame per slot");
    CU(c, cudaSetDevice(c->device));
    const size_t slot_bytes = (size_t)frames_per_slot * c->touched.size() * std::max(c->row_bytes, c->tap_row_bytes);
    CU(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CU(c, cudaStreamCreateWithFlags(&c->compute_stream, cudaStreamNonBlocking));
    c->ring.resize(n_slots);
    c->frames_per_slot = frames_per_slot;
    c->next_slot = 0;
    for (auto& s : c->ring) {
        // the pinned staging buffer is only needed for pageable sources: allocated on first use
        CU(c, esdguard::gmalloc(&s.d_rows, slot_bytes));
        CU(c, cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming));
        CU(c, cudaEventCreateWithFlags(&s.consumed, cudaEventDisableTiming));
    }
    return ESD_OK;
}

int esd_ingest_close(esd_ctx* c) {
    if (!c) return ESD_ERR_INVALID;
    if (c->ring.empty()) return ESD_OK;
    cudaSetDevice(c->device);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    if (c->compute_stream) cudaStreamSynchronize(c->compute_stream);
    for (auto& s : c->ring) {
        if (s.h_pinned) cudaFreeHost(s.h_pinned);
        esdguard::gfree(s.d_rows);
        if (s.copied) cudaEventDestroy(s.copied);
        if (s.consumed) cudaEventDestroy(s.consumed);
    }
    c->ring.clear();
    if (c->have_last_stream && (c->last_stream == c->compute_stream)) c->have_last_stream = false;
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->compute_stream) cudaStreamDestroy(c->compute_stream);
    c->copy_stream = c->compute_stream = nullptr;
    return ESD_OK;
}

int esd_ingest_push_host(esd_ctx* c, const uint8_t* h_bgr, int64_t n, int64_t frame_stride, int64_t pitch,
                         int64_t first_frame_num) {
    if (!c) return ESD_ERR_INVALID;
    if (c->ring.empty()) return fail(c, ESD_ERR_STATE, "ingest ring not open");
    if (!h_bgr || n <= 0) return fail(c, ESD_ERR_INVALID, "ingest: null frames or n <= 0");
    if (pitch < c->row_bytes) return fail(c, ESD_ERR_INVALID, "ingest: pitch < row bytes");
    if (c->i420 && (pitch & 1)) return fail(c, ESD_ERR_INVALID, "ingest: I420 frames need an even pitch (chroma rows are pitch / 2 apart)");
    CU(c, cudaSetDevice(c->device));
    cudaPointerAttributes attr;
    bool pinned = false;
    if (cudaPointerGetAttributes(&attr, h_bgr) == cudaSuccess) pinned = (attr.type == cudaMemoryTypeHost);
    else cudaGetLastError();
    const int64_t nt = (int64_t)c->touched.size();
    const int64_t cfb = nt * c->row_bytes;  // compact frame bytes
    // runs of consecutive touched rows: (first compact index, first source row, length)
    struct Run { int crow, row, len; };
    std::vector<Run> runs;
    // I420: only the Y rows go by runs; a touched chroma row is two half rows (U, V) from two planes, copied side by side into its
    // compact row and interleaved on the device (i420_interleave_kernel, in place) before the fused kernel reads it
    const int nt_runs = c->i420 ? c->n_touched_y : (int)nt;
    const int H = c->cfg.src_height, half_row = c->row_bytes / 2;
    const int64_t u_off = (int64_t)H * pitch, v_off = u_off + (int64_t)(H / 2) * (pitch / 2);
    for (int i = 0; i < nt_runs;) {
        int j = i + 1;
        while (j < nt_runs && c->touched[j] == c->touched[j - 1] + 1 && pitch == c->row_bytes) ++j;
        runs.push_back(Run{i, c->touched[i], j - i});
        i = j;
    }
    // the tap gather only pays when it shrinks the rows (downscale factor > 2)
    const bool gather = c->gather_threads > 0 && c->resize && c->tap_row_bytes < c->row_bytes;
    const int64_t tfb = nt * c->tap_row_bytes;  // tap-compact frame bytes
    for (int64_t done = 0; done < n;) {
        const int64_t m = std::min<int64_t>(c->frames_per_slot, n - done);
        IngestSlot& s = c->ring[c->next_slot];
        c->next_slot = (c->next_slot + 1) % (int)c->ring.size();
        TraceTimer trw;
        if (s.in_flight) CU(c, cudaEventSynchronize(s.consumed));  // device slot (and pinned slot) free again
        trw.lap("wait slot");
        const uint8_t* src = h_bgr + done * frame_stride;
        // (A per-slot DMA + gather hybrid was measured twice and is slower: both share PCIe and the in-order pushes stall
        // behind the 4x larger DMA slots -- 31-36 k vs 55 k frames/s, profiles/r01_gather_prefetch.log.)
        const bool use_gather = gather;
        if (use_gather) {
            // host threads gather, per touched row, the two BGR taps of every destination column (6 of every
            // ~3*scale bytes) into the pinned slot: 3.75x fewer PCIe bytes again at 1080p (442 KB per frame)
            { int rcp = ensure_pinned(c, s, (size_t)c->frames_per_slot * tfb); if (rcp) return rcp; }
            static const int pf_env = getenv("ESD_GATHER_PF") ? atoi(getenv("ESD_GATHER_PF")) : -1;
            static const bool nt_off = getenv("ESD_GATHER_NT") && atoi(getenv("ESD_GATHER_NT")) == 0;  // A/B switch
            GatherSpec gs{};
            gs.dst_w = c->dst_w;
            gs.tap_row_bytes = c->tap_row_bytes;
            gs.row_bytes = c->row_bytes;
            gs.off = c->tap_src_off.data();
            gs.touched = c->touched.data();
            gs.n_touched = nt;
            gs.n_touched_y = c->n_touched_y;
            gs.nv12 = c->nv12;
            gs.i420 = c->i420;
            gs.src_height = H;
            // measured best of 0 / 512 ... 8192 at 1080p (profiles/r01_gather_prefetch.log); never further than one row
            gs.prefetch_bytes = std::min(pf_env >= 0 ? pf_env : 4096, c->row_bytes);
            gs.nt_stores = !nt_off;  // +11 % with 16 threads (profiles/r01_gather_prefetch.log)
            static const int streams_env = getenv("ESD_GATHER_STREAMS") ? atoi(getenv("ESD_GATHER_STREAMS")) : kGatherStreams;  // A/B switch: 1
            static const int pf8_env = getenv("ESD_GATHER_PF8") ? atoi(getenv("ESD_GATHER_PF8")) : 384;
            gs.streams = streams_env;
            gs.prefetch_bytes_multi = pf8_env;
            gs.prefetch_hint = 0;
            uint8_t* dst_base = s.h_pinned;
            std::function<void(int64_t, int64_t)> job = [=](int64_t lo, int64_t hi) {
                gather_tap_rows(gs, src, frame_stride, pitch, dst_base, lo, hi);
            };
            TraceTimer tr;
            SharedGatherPool::instance().parallel_for(m * nt, 16, c->gather_threads, job);
            tr.lap("host tap gather");
            CU(c, cudaMemcpyAsync(s.d_rows, s.h_pinned, (size_t)(m * tfb), cudaMemcpyHostToDevice, c->copy_stream));
            tr.lap("memcpyAsync call");
            c->h2d_copies++;
            c->h2d_bytes += m * tfb;
            CU(c, cudaEventRecord(s.copied, c->copy_stream));
            CU(c, cudaStreamWaitEvent(c->compute_stream, s.copied, 0));
            int rc = push_common(c, s.d_rows, m, tfb, c->tap_row_bytes, LAYOUT_TAPS, first_frame_num + done, c->compute_stream);
            if (rc) return rc;
            CU(c, cudaEventRecord(s.consumed, c->compute_stream));
            s.in_flight = true;
            done += m;
            continue;
        }
        if (pinned) {
            // DMA straight from the caller's pinned frames: one strided 2-D copy per row run
            // (rows = frames), so only touched rows cross PCIe.
            for (const Run& r : runs) {
                CU(c, cudaMemcpy2DAsync(s.d_rows + (int64_t)r.crow * c->row_bytes, (size_t)cfb,
                                        src + (int64_t)r.row * pitch, (size_t)frame_stride,
                                        (size_t)r.len * c->row_bytes, (size_t)m, cudaMemcpyHostToDevice, c->copy_stream));
                c->h2d_copies++;
            }
            for (int i = nt_runs; i < (int)nt; ++i) {
                const int64_t r = c->touched[i] - H;
                uint8_t* drow = s.d_rows + (int64_t)i * c->row_bytes;
                CU(c, cudaMemcpy2DAsync(drow, (size_t)cfb, src + u_off + r * (pitch / 2), (size_t)frame_stride, (size_t)half_row, (size_t)m,
                                        cudaMemcpyHostToDevice, c->copy_stream));
                CU(c, cudaMemcpy2DAsync(drow + half_row, (size_t)cfb, src + v_off + r * (pitch / 2), (size_t)frame_stride, (size_t)half_row,
                                        (size_t)m, cudaMemcpyHostToDevice, c->copy_stream));
                c->h2d_copies += 2;
            }
        } else {
            // pageable source: the CPU gathers the touched rows into the pinned slot
            { int rcp = ensure_pinned(c, s, (size_t)c->frames_per_slot * cfb); if (rcp) return rcp; }
            for (int64_t f = 0; f < m; ++f)
                for (const Run& r : runs)
                    memcpy(s.h_pinned + f * cfb + (int64_t)r.crow * c->row_bytes, src + f * frame_stride + (int64_t)r.row * pitch,
                           (size_t)r.len * c->row_bytes);
            for (int64_t f = 0; f < m; ++f)
                for (int i = nt_runs; i < (int)nt; ++i) {
                    const int64_t r = c->touched[i] - H;
                    uint8_t* drow = s.h_pinned + f * cfb + (int64_t)i * c->row_bytes;
                    memcpy(drow, src + f * frame_stride + u_off + r * (pitch / 2), (size_t)half_row);
                    memcpy(drow + half_row, src + f * frame_stride + v_off + r * (pitch / 2), (size_t)half_row);
                }
            CU(c, cudaMemcpyAsync(s.d_rows, s.h_pinned, (size_t)(m * cfb), cudaMemcpyHostToDevice, c->copy_stream));
            c->h2d_copies++;
        }
        c->h2d_bytes += m * cfb;
        CU(c, cudaEventRecord(s.copied, c->copy_stream));
        CU(c, cudaStreamWaitEvent(c->compute_stream, s.copied, 0));
        if (c->i420) {
            int rci = order_after_last(c, c->compute_stream);
            uint8_t* uv0 = s.d_rows + (int64_t)c->n_touched_y * c->row_bytes;
            if (!rci) rci = interleave_chroma(c, uv0, uv0 + half_row, cfb, c->row_bytes, nullptr, uv0, cfb, c->row_bytes, m, c->compute_stream);
            if (rci) return rci;
        }
        int rc = esd_push_rows(c, s.d_rows, m, first_frame_num + done, c->compute_stream);
        if (rc) return rc;
        CU(c, cudaEventRecord(s.consumed, c->compute_stream));
        s.in_flight = true;
        done += m;
    }
    return ESD_OK;
}

int esd_ingest_set_gather(esd_ctx* c, int32_t n_threads) {
    if (!c) return ESD_ERR_INVALID;
    if (n_threads < 0 || n_threads > 256) return fail(c, ESD_ERR_INVALID, "ingest: gather threads must be in 0..256");
    if (n_threads > 0 && !c->resize) return fail(c, ESD_ERR_UNSUPPORTED, "ingest: tap gather needs a resizing context");
    CU(c, cudaSetDevice(c->device));
    int rc = sync_all(c);
    if (rc) return rc;
    // n_threads workers of the process-wide pool at most (the pool itself never exceeds the node budget)
    c->gather_threads = n_threads > 0 ? std::min(n_threads, SharedGatherPool::instance().ensure(n_threads)) : 0;
    return ESD_OK;
}

int esd_ingest_wait_copied(esd_ctx* c) {
    if (!c) return ESD_ERR_INVALID;
    if (c->ring.empty()) return fail(c, ESD_ERR_STATE, "ingest ring not open");
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->copy_stream));  // the H2D copies only: scoring keeps running on the compute stream
    return ESD_OK;
}

int esd_ingest_stats(const esd_ctx* c, int64_t* bytes, int64_t* copies) {
    if (!c) return ESD_ERR_INVALID;
    if (bytes) *bytes = c->h2d_bytes;
    if (copies) *copies = c->h2d_copies;
    return ESD_OK;
}

int esd_synchronize(esd_ctx* c) {
    if (!c) return ESD_ERR_INVALID;
    CU(c, cudaSetDevice(c->device));
    return sync_all(c);
}

int esd_join(esd_ctx* c, void* stream) {
    if (!c) return ESD_ERR_INVALID;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaEventRecord(c->ev_join, c->aux_stream));
    CU(c, cudaStreamWaitEvent((cudaStream_t)stream, c->ev_join, 0));
    return ESD_OK;
}

int64_t esd_frames_pushed(const esd_ctx* c) { return c ? c->n_frames : 0; }

int esd_read_scores(esd_ctx* c, int64_t from_frame, int64_t n, uint64_t* sums3, double* content_val,
                    double* adaptive_val, double* adaptive_ratio, uint32_t* hist, double* hist_diff) {
    if (!c) return ESD_ERR_INVALID;
    if (n == 0) return ESD_OK;
    const int64_t i0 = from_frame - c->first_frame;
    if (!c->started || n < 0 || i0 < 0 || i0 + n > c->n_frames)
        return fail(c, ESD_ERR_INVALID, "read_scores: range [%lld, %lld) outside pushed frames", (long long)from_frame,
                    (long long)(from_frame + n));
    int rc = esd_synchronize(c);
    if (rc) return rc;
    if ((sums3 || content_val || adaptive_val || adaptive_ratio) && !c->need_content)
        return fail(c, ESD_ERR_STATE, "read_scores: content scores requested but no content/adaptive detector configured");
    if ((hist || hist_diff) && !c->need_hist)
        return fail(c, ESD_ERR_STATE, "read_scores: histogram requested but no histogram detector configured");
    if (sums3) CU(c, cudaMemcpy(sums3, c->d_sums3 + 3 * i0, sizeof(uint64_t) * 3 * n, cudaMemcpyDeviceToHost));
    if (content_val) CU(c, cudaMemcpy(content_val, c->d_cv + i0, sizeof(double) * n, cudaMemcpyDeviceToHost));
    if (adaptive_val) CU(c, cudaMemcpy(adaptive_val, c->d_av + i0, sizeof(double) * n, cudaMemcpyDeviceToHost));
    if (adaptive_ratio) CU(c, cudaMemcpy(adaptive_ratio, c->d_ratio + i0, sizeof(double) * n, cudaMemcpyDeviceToHost));
    if (hist) CU(c, cudaMemcpy(hist, c->d_counts + i0 * c->cfg.hist_bins, sizeof(uint32_t) * n * c->cfg.hist_bins, cudaMemcpyDeviceToHost));
    if (hist_diff) CU(c, cudaMemcpy(hist_diff, c->d_hdiff + i0, sizeof(double) * n, cudaMemcpyDeviceToHost));
    return ESD_OK;
}

static int det_index(int32_t detector) {
    return detector == ESD_DET_CONTENT ? 0 : detector == ESD_DET_ADAPTIVE ? 1 : detector == ESD_DET_HIST ? 2
         : detector == ESD_DET_THRESHOLD ? 3 : detector == ESD_DET_HASH ? 4 : -1;
}

int esd_read_edge_counts(esd_ctx* c, int64_t from_frame, int64_t n, uint32_t* counts) {
    if (!c || !counts) return ESD_ERR_INVALID;
    if (n == 0) return ESD_OK;
    const int64_t i0 = from_frame - c->first_frame;
    if (!c->started || n < 0 || i0 < 0 || i0 + n > c->n_frames)
        return fail(c, ESD_ERR_INVALID, "read_edge_counts: range outside pushed frames");
    if (!c->need_edges) return fail(c, ESD_ERR_STATE, "read_edge_counts: no delta_edges weight configured");
    int rc = esd_synchronize(c);
    if (rc) return rc;
    CU(c, cudaMemcpy(counts, c->d_edge_counts + i0, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
    return ESD_OK;
}

int esd_read_hash(esd_ctx* c, int64_t from_frame, int64_t n, uint32_t* bits, double* hash_dist) {
    if (!c) return ESD_ERR_INVALID;
    if (n == 0) return ESD_OK;
    const int64_t i0 = from_frame - c->first_frame;
    if (!c->started || n < 0 || i0 < 0 || i0 + n > c->n_frames)
        return fail(c, ESD_ERR_INVALID, "read_hash: range outside pushed frames");
    if (!c->need_hash) return fail(c, ESD_ERR_STATE, "read_hash: no hash detector configured");
    int rc = esd_synchronize(c);
    if (rc) return rc;
    if (bits) CU(c, cudaMemcpy(bits, c->d_hash + i0 * c->hash_words, sizeof(uint32_t) * n * c->hash_words, cudaMemcpyDeviceToHost));
    if (hash_dist) CU(c, cudaMemcpy(hash_dist, c->d_hdist + i0, sizeof(double) * n, cudaMemcpyDeviceToHost));
    return ESD_OK;
}

int esd_read_hash_margin(esd_ctx* c, int64_t from_frame, int64_t n, float* min_margin) {
    if (!c || !min_margin) return ESD_ERR_INVALID;
    if (n == 0) return ESD_OK;
    const int64_t i0 = from_frame - c->first_frame;
    if (!c->started || n < 0 || i0 < 0 || i0 + n > c->n_frames)
        return fail(c, ESD_ERR_INVALID, "read_hash_margin: range outside pushed frames");
    if (!c->need_hash) return fail(c, ESD_ERR_STATE, "read_hash_margin: no hash detector configured");
    int rc = esd_synchronize(c);
    if (rc) return rc;
    CU(c, cudaMemcpy(min_margin, c->d_hmargin + i0, sizeof(float) * n, cudaMemcpyDeviceToHost));
    return ESD_OK;
}

int esd_debug_read_hash_input(esd_ctx* c, int64_t frame, uint8_t* out, int64_t cap) {
    if (!c || !out) return ESD_ERR_INVALID;
    if (!c->need_hash) return fail(c, ESD_ERR_STATE, "debug_read_hash_input: no hash detector configured");
    const int64_t i = frame - c->first_frame - c->last_batch_base;
    if (!c->started || i < 0 || i >= c->last_batch_n)
        return fail(c, ESD_ERR_INVALID, "debug_read_hash_input: frame %lld is not part of the most recent push", (long long)frame);
    const int64_t sz = (int64_t)c->hparams.S * c->hparams.S;
    if (cap < sz) return ESD_ERR_CAPACITY;
    int rc = esd_synchronize(c);
    if (rc) return rc;
    CU(c, cudaMemcpy(out, c->d_hash_small + i * sz, (size_t)sz, cudaMemcpyDeviceToHost));
    return ESD_OK;
}

int esd_read_average_rgb(esd_ctx* c, int64_t from_frame, int64_t n, double* average_rgb) {
    if (!c || !average_rgb) return ESD_ERR_INVALID;
    if (n == 0) return ESD_OK;
    const int64_t i0 = from_frame - c->first_frame;
    if (!c->started || n < 0 || i0 < 0 || i0 + n > c->n_frames)
        return fail(c, ESD_ERR_INVALID, "read_average_rgb: range outside pushed frames");
    if (!(c->cfg.detectors & ESD_DET_THRESHOLD)) return fail(c, ESD_ERR_STATE, "read_average_rgb: no threshold detector configured");
    int rc = esd_synchronize(c);
    if (rc) return rc;
    CU(c, cudaMemcpy(average_rgb, c->d_avg + i0, sizeof(double) * n, cudaMemcpyDeviceToHost));
    return ESD_OK;
}

int esd_post_process(esd_ctx* c, int32_t detector, int64_t last_frame_num, int64_t* cuts, int64_t cap, int64_t* n_cuts) {
    if (!c) return ESD_ERR_INVALID;
    if (det_index(detector) < 0 || !(c->cfg.detectors & detector)) return fail(c, ESD_ERR_INVALID, "post_process: detector %d not configured", detector);
    if (n_cuts) *n_cuts = 0;
    if (detector != ESD_DET_THRESHOLD) return ESD_OK;  // the other detectors' post_process returns []
    int rc = esd_synchronize(c);
    if (rc) return rc;
    DecisionState st;
    CU(c, cudaMemcpy(&st, c->d_state, sizeof st, cudaMemcpyDeviceToHost));
    // ThresholdDetector.post_process: close the scene at the last fade-out when asked to
    const int L = c->cfg.thresh_min_scene_len;
    const bool len_ok = st.t_init ? (last_frame_num - st.t_last_scene_cut) >= L : last_frame_num >= L;
    if (st.t_processed && st.t_fade_out && c->cfg.thresh_add_final_scene && len_ok) {
        if (cap < 1 || !cuts) return fail(c, ESD_ERR_CAPACITY, "post_process: buffer too small");
        cuts[0] = st.t_fade_frame;
        if (n_cuts) *n_cuts = 1;
    }
    return ESD_OK;
}

int esd_get_cuts(esd_ctx* c, int32_t detector, int64_t from_index, int64_t* cuts, int64_t cap, int64_t* n_written,
                 int64_t* n_total) {
    if (!c) return ESD_ERR_INVALID;
    const int di = det_index(detector);
    if (di < 0 || !(c->cfg.detectors & detector)) return fail(c, ESD_ERR_INVALID, "get_cuts: detector %d not configured", detector);
    int rc = esd_synchronize(c);
    if (rc) return rc;
    DecisionState st;
    CU(c, cudaMemcpy(&st, c->d_state, sizeof st, cudaMemcpyDeviceToHost));
    if (st.overflow) return fail(c, ESD_ERR_CAPACITY, "cut list overflow (max_cuts = %lld)", (long long)c->max_cuts);
    const int64_t total = st.n_cuts[di];
    if (n_total) *n_total = total;
    if (from_index < 0) from_index = 0;
    const int64_t avail = std::max<int64_t>(0, total - from_index);
    const int64_t m = std::min(avail, cap);
    if (n_written) *n_written = m;
    if (m > 0 && cuts)
        CU(c, cudaMemcpy(cuts, c->d_cuts + (int64_t)di * c->max_cuts + from_index, sizeof(int64_t) * m, cudaMemcpyDeviceToHost));
    if (avail > cap) return fail(c, ESD_ERR_CAPACITY, "get_cuts: %lld cuts pending, buffer holds %lld", (long long)avail, (long long)cap);
    return ESD_OK;
}

int esd_process_frame_host(esd_ctx* c, const uint8_t* h_bgr, int64_t pitch, int64_t frame_num, int32_t detector,
                           int64_t from_index, int64_t* cuts, int64_t cap, int64_t* n_written, int64_t* n_total) {
    if (!c) return ESD_ERR_INVALID;
    const int di = det_index(detector);
    if (di < 0 || !(c->cfg.detectors & detector)) return fail(c, ESD_ERR_INVALID, "process_frame: detector %d not configured", detector);
    if (!h_bgr) return fail(c, ESD_ERR_INVALID, "process_frame: null frame");
    if (c->nv12) return fail(c, ESD_ERR_UNSUPPORTED, "process_frame: BGR24 contexts only");
    if (pitch < c->row_bytes) return fail(c, ESD_ERR_INVALID, "process_frame: pitch %lld < row bytes %d", (long long)pitch, c->row_bytes);
    CU(c, cudaSetDevice(c->device));
    const int H = c->cfg.src_height;
    const size_t fb = (size_t)H * c->row_bytes;
    // first call: each resource on its own, so a failed allocation is retried instead of leaving a half-built path
    if (!c->pf_stream) CU(c, cudaStreamCreateWithFlags(&c->pf_stream, cudaStreamNonBlocking));
    if (!c->pf_h) CU(c, cudaHostAlloc(&c->pf_h, fb + 64, cudaHostAllocMapped));  // + slack: TMA copies are rounded up to 16 bytes
    if (!c->pf_mailbox) {
        CU(c, cudaHostAlloc(&c->pf_mailbox, 16 * sizeof(long long), cudaHostAllocMapped));
        memset(c->pf_mailbox, 0, 16 * sizeof(long long));
    }
    cudaStream_t st = c->pf_stream;
    // work enqueued through another entry point may still be running its tail on the library stream
    if (c->have_last_stream && c->last_stream != st) CU(c, cudaStreamSynchronize(c->aux_stream));
    TraceTimer trp;
    if (pitch == c->row_bytes) memcpy(c->pf_h, h_bgr, fb);
    else for (int y = 0; y < H; ++y) memcpy(c->pf_h + (size_t)y * c->row_bytes, h_bgr + (size_t)y * pitch, (size_t)c->row_bytes);
    // zero-copy: the fused kernel reads the pinned frame over PCIe (110 KB at detector resolution) and the decision
    // kernel posts the cut counts into host-mapped memory, so the call is three kernel launches and one synchronise
    trp.lap("pf stage memcpy");
    c->h2d_bytes += (int64_t)fb;
    c->h2d_copies++;
    // The kernel sequence of a one-frame push is fixed once the adaptive window has filled, so it runs as a CUDA graph
    // whose node parameters are refreshed per call (ESD_NO_GRAPH=1 keeps plain launches, for A/B timing).
    static const bool no_graph = getenv("ESD_NO_GRAPH") != nullptr;
    const int64_t warm = (c->cfg.detectors & ESD_DET_ADAPTIVE) ? 2LL * c->cfg.adaptive_window_width + 1 : 1;
    const bool use_graph = !no_graph && !c->need_edges && !c->timing && c->started && c->n_frames >= warm;
    KernelGraph& kg = c->kg;
    int rc;
    ++c->pf_ticket;
    if (use_graph) {
        if (!kg.exec) {
            kg.destroy();
            CU(c, cudaGraphCreate(&kg.graph, 0));
            kg.mode = KernelGraph::BUILD;
        } else {
            kg.mode = KernelGraph::UPDATE;
            kg.cursor = 0;
        }
        rc = push_common(c, c->pf_h, 1, (int64_t)fb, c->row_bytes, LAYOUT_FULL, frame_num, st, true, c->pf_mailbox);
        const int mode = kg.mode;
        kg.mode = KernelGraph::DIRECT;
        if (rc) { kg.destroy(); return rc; }
        if (mode == KernelGraph::BUILD) CU(c, cudaGraphInstantiate(&kg.exec, kg.graph, 0));
        CU(c, cudaGraphLaunch(kg.exec, st));
    } else {
        rc = push_common(c, c->pf_h, 1, (int64_t)fb, c->row_bytes, LAYOUT_FULL, frame_num, st, true, c->pf_mailbox);
        if (rc) return rc;
    }
    trp.lap("pf enqueue");
    {   // poll the completion tickets of the configured detectors in host-mapped memory (the kernels are a few tens of
        // microseconds); a stuck or faulted launch falls through to cudaStreamSynchronize, which reports the error
        volatile long long* mb = c->pf_mailbox;
        const double t_poll = TraceTimer::now();
        bool done = false;
        for (uint32_t spins = 0; !done; ++spins) {
            done = true;
            for (int d = 0; d < 5; ++d)
                if ((c->cfg.detectors & (1 << d)) && mb[8 + d] != c->pf_ticket) { done = false; break; }
            if (!done) {
                _mm_pause();
                if ((spins & 0xfff) == 0xfff && TraceTimer::now() - t_poll > 20.0) break;  // 20 ms: give up polling
            }
        }
        if (!done) CU(c, cudaStreamSynchronize(st));
    }
    trp.lap("pf wait");
    if (c->pf_mailbox[5]) return fail(c, ESD_ERR_CAPACITY, "cut list overflow (max_cuts = %lld)", (long long)c->max_cuts);
    const int64_t total = c->pf_mailbox[di];
    if (n_total) *n_total = total;
    if (from_index < 0) from_index = 0;
    const int64_t avail = std::max<int64_t>(0, total - from_index);
    const int64_t m = std::min(avail, cap);
    if (n_written) *n_written = m;
    if (m > 0 && cuts) {
        CU(c, cudaStreamSynchronize(st));  // the cut list itself lives in device memory
        CU(c, cudaMemcpy(cuts, c->d_cuts + (int64_t)di * c->max_cuts + from_index, sizeof(int64_t) * m, cudaMemcpyDeviceToHost));
    }
    if (avail > cap) return fail(c, ESD_ERR_CAPACITY, "process_frame: %lld cuts pending, buffer holds %lld", (long long)avail, (long long)cap);
    return ESD_OK;
}

// Scratch of the stand-alone decision pass: device score/ratio arrays and the state block grow only; results (cut
// counts, completion ticket, the cuts themselves) are written by decide_kernel straight into host-mapped pinned memory.
static int ensure_decide_scratch(esd_ctx* c, int64_t n_dev, int64_t n_host_stage) {
    if (!c->dec_stream) CU(c, cudaStreamCreateWithFlags(&c->dec_stream, cudaStreamNonBlocking));
    if (!c->dec_d_state) CU(c, esdguard::gmalloc(&c->dec_d_state, sizeof(DecisionState)));
    if (n_dev > c->dec_cap) {
        const int64_t ncap = std::max<int64_t>(n_dev, std::max<int64_t>(32768, 2 * c->dec_cap));
        CU(c, cudaStreamSynchronize(c->dec_stream));
        esdguard::gfree(c->dec_d_scores); esdguard::gfree(c->dec_d_ratio);
        c->dec_d_scores = c->dec_d_ratio = nullptr;
        c->dec_cap = 0;
        CU(c, esdguard::gmalloc(&c->dec_d_scores, sizeof(double) * ncap));
        CU(c, esdguard::gmalloc(&c->dec_d_ratio, sizeof(double) * ncap));
        c->dec_cap = ncap;
    }
    if (!c->dec_h || n_host_stage > c->dec_h_scores) {
        const int64_t ns = std::max<int64_t>(n_host_stage, std::max<int64_t>(32768, 2 * c->dec_h_scores));
        if (c->dec_h) { CU(c, cudaStreamSynchronize(c->dec_stream)); cudaFreeHost(c->dec_h); c->dec_h = nullptr; c->dec_h_scores = 0; }
        CU(c, cudaHostAlloc(&c->dec_h, 16 * sizeof(long long) + sizeof(long long) * c->max_cuts + sizeof(double) * ns, cudaHostAllocMapped));
        memset(c->dec_h, 0, 16 * sizeof(long long));
        c->dec_h_scores = ns;
    }
    return ESD_OK;
}

// Enqueue the pass over device-resident scores on `st`, wait for this pass only (ticket poll, stream sync as fallback) and
// hand the cuts back.  d_ratio: device array [n] that receives the adaptive ratios (ESD_DET_ADAPTIVE; library scratch if null).
static int decide_on_device(esd_ctx* c, int32_t detector, int di, int64_t first_frame_num, int64_t n, const double* d_scores,
                            double* d_ratio, cudaStream_t st, int64_t* cuts, int64_t cap, int64_t* n_cuts, bool sync_stream) {
    long long* mailbox = reinterpret_cast<long long*>(c->dec_h);
    long long* h_cuts = mailbox + 16;
    mailbox[di] = 0; mailbox[5] = 0;
    const long long ticket = ++c->dec_ticket;
    DecisionParams P = c->dparams;
    P.detectors = detector;
    P.cuts_stride = 0;  // one list: only `detector` runs
    P.fresh_state = 1;  // starts from the zero state inside the kernel: no memset launch, `dec_d_state` is never touched
    if (detector == ESD_DET_ADAPTIVE && d_ratio) {  // the caller wants the ratios: a pass of their own; otherwise computed inline
        adaptive_ratio_full_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_scores, d_ratio, (long long)n, P.adaptive_w,
                                                                                P.adaptive_min_content_val);
        CU(c, cudaGetLastError());
        c->launches++;
    }
    decide_kernel<<<5, kDecideThreads, 0, st>>>(P, c->dec_d_state, h_cuts, d_scores, d_scores, d_ratio, d_scores, d_scores, d_scores,
                                                first_frame_num, 0, n, mailbox, ticket);
    CU(c, cudaGetLastError());
    c->launches++;
    if (sync_stream) {
        CU(c, cudaStreamSynchronize(st));
    } else {
        volatile long long* mb = mailbox;
        const double t_poll = TraceTimer::now();
        bool done = false;
        for (uint32_t spins = 0; !(done = (mb[8 + di] == ticket)); ++spins) {
            _mm_pause();
            if ((spins & 0xfff) == 0xfff && TraceTimer::now() - t_poll > 50.0) break;  // 50 ms: let the runtime report what happened
        }
        if (!done) CU(c, cudaStreamSynchronize(st));
    }
    const int64_t total = mailbox[di];
    if (n_cuts) *n_cuts = total;
    if (mailbox[5] || total > cap) return fail(c, ESD_ERR_CAPACITY, "decide: %lld cuts, buffer holds %lld (max_cuts %lld)", (long long)total, (long long)cap, (long long)c->max_cuts);
    if (total > 0 && cuts) memcpy(cuts, h_cuts, sizeof(int64_t) * total);
    return ESD_OK;
}

int esd_decide_arrays(esd_ctx* c, int32_t detector, int64_t first_frame_num, int64_t n, const double* scores,
                      double* adaptive_ratio_out, int64_t* cuts, int64_t cap, int64_t* n_cuts) {
    if (!c || !scores || n < 0) return ESD_ERR_INVALID;
    const int di = det_index(detector);
    if (di < 0 || !(c->cfg.detectors & detector)) return fail(c, ESD_ERR_INVALID, "decide_arrays: detector %d not configured", detector);
    CU(c, cudaSetDevice(c->device));
    if (n_cuts) *n_cuts = 0;
    if (n == 0) return ESD_OK;
    int rc = ensure_decide_scratch(c, n, n);
    if (rc) return rc;
    double* stage = reinterpret_cast<double*>(c->dec_h + 16 * sizeof(long long) + sizeof(long long) * c->max_cuts);
    memcpy(stage, scores, sizeof(double) * n);
    CU(c, cudaMemcpyAsync(c->dec_d_scores, stage, sizeof(double) * n, cudaMemcpyHostToDevice, c->dec_stream));
    if (!(adaptive_ratio_out && detector == ESD_DET_ADAPTIVE)) {
        rc = decide_on_device(c, detector, di, first_frame_num, n, c->dec_d_scores, nullptr, c->dec_stream, cuts, cap, n_cuts, false);
        if (adaptive_ratio_out) for (int64_t i = 0; i < n; ++i) adaptive_ratio_out[i] = NAN;  // no ratios for the other detectors
        return rc;
    }
    // adaptive with the ratios wanted on the host: the pass, then the D2H copy of the ratios behind it
    rc = decide_on_device(c, detector, di, first_frame_num, n, c->dec_d_scores, c->dec_d_ratio, c->dec_stream, cuts, cap, n_cuts, false);
    if (rc && rc != ESD_ERR_CAPACITY) return rc;
    CU(c, cudaMemcpyAsync(stage, c->dec_d_ratio, sizeof(double) * n, cudaMemcpyDeviceToHost, c->dec_stream));
    CU(c, cudaStreamSynchronize(c->dec_stream));
    memcpy(adaptive_ratio_out, stage, sizeof(double) * n);
    return rc;
}

int esd_decide_device(esd_ctx* c, int32_t detector, int64_t first_frame_num, int64_t n, const double* d_scores,
                      double* d_adaptive_ratio_out, int64_t* cuts, int64_t cap, int64_t* n_cuts, void* stream) {
    if (!c || !d_scores || n < 0) return ESD_ERR_INVALID;
    const int di = det_index(detector);
    if (di < 0 || !(c->cfg.detectors & detector)) return fail(c, ESD_ERR_INVALID, "decide_device: detector %d not configured", detector);
    CU(c, cudaSetDevice(c->device));
    if (n_cuts) *n_cuts = 0;
    if (n == 0) return ESD_OK;
    int rc = validate_device_span(c, reinterpret_cast<const uint8_t*>(d_scores), sizeof(double) * (size_t)n, "decide_device (scores)");
    if (rc) return rc;
    if (d_adaptive_ratio_out) {
        rc = validate_device_span(c, reinterpret_cast<const uint8_t*>(d_adaptive_ratio_out), sizeof(double) * (size_t)n, "decide_device (ratios)");
        if (rc) return rc;
    }
    rc = ensure_decide_scratch(c, 0, 0);
    if (rc) return rc;
    return decide_on_device(c, detector, di, first_frame_num, n, d_scores, d_adaptive_ratio_out, (cudaStream_t)stream, cuts, cap, n_cuts, false);
}

int esd_copy_scores_device(esd_ctx* c, int32_t kind, int64_t from_frame, int64_t n, double* d_dst, int32_t dst_device, void* stream) {
    if (!c || !d_dst || n < 0) return ESD_ERR_INVALID;
    if (n == 0) return ESD_OK;
    const int64_t i0 = from_frame - c->first_frame;
    if (!c->started || i0 < 0 || i0 + n > c->n_frames)
        return fail(c, ESD_ERR_INVALID, "copy_scores_device: range [%lld, %lld) outside pushed frames", (long long)from_frame, (long long)(from_frame + n));
    const double* src = nullptr;
    switch (kind) {
        case ESD_SCORE_CONTENT_VAL: src = c->need_content ? c->d_cv : nullptr; break;
        case ESD_SCORE_ADAPTIVE_VAL: src = c->need_content ? c->d_av : nullptr; break;
        case ESD_SCORE_ADAPTIVE_RATIO: src = (c->cfg.detectors & ESD_DET_ADAPTIVE) ? c->d_ratio : nullptr; break;
        case ESD_SCORE_HIST_DIFF: src = c->need_hist ? c->d_hdiff : nullptr; break;
        case ESD_SCORE_AVERAGE_RGB: src = (c->cfg.detectors & ESD_DET_THRESHOLD) ? c->d_avg : nullptr; break;
        case ESD_SCORE_HASH_DIST: src = c->need_hash ? c->d_hdist : nullptr; break;
        default: return fail(c, ESD_ERR_INVALID, "copy_scores_device: unknown score kind %d", kind);
    }
    if (!src) return fail(c, ESD_ERR_STATE, "copy_scores_device: score kind %d is not produced by the configured detectors", kind);
    CU(c, cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    // order the copy behind every finalize tail enqueued so far (they run on the library's stream)
    CU(c, cudaEventRecord(c->ev_join, c->aux_stream));
    CU(c, cudaStreamWaitEvent(st, c->ev_join, 0));
    if (dst_device == c->device || dst_device < 0)
        CU(c, cudaMemcpyAsync(d_dst, src + i0, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    else
        CU(c, cudaMemcpyPeerAsync(d_dst, dst_device, src + i0, c->device, sizeof(double) * n, st));
    return ESD_OK;
}

int esd_debug_read_prev(esd_ctx* c, uint32_t* out, int64_t cap) {
    if (!c || !out) return ESD_ERR_INVALID;
    if (!c->need_content || c->n_frames == 0) return fail(c, ESD_ERR_STATE, "debug_read_prev: no content frame pushed yet");
    const int64_t n = (int64_t)c->dst_w * c->dst_h;
    if (cap < n) return ESD_ERR_CAPACITY;
    int rc = esd_synchronize(c);
    if (rc) return rc;
    CU(c, cudaMemcpy(out, c->d_prev[c->prev_parity], sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
    return ESD_OK;
}

int esd_debug_guard_selftest(int32_t damage) {
    // with ESD_GUARD=1: allocates a guarded buffer, optionally writes one byte past its end, frees it (the free verifies the
    // red zones: a damaged one aborts the process).  Returns 1 when the guards are active, 0 when the wrappers are plain.
    if (!esdguard::enabled()) return 0;
    uint8_t* p = nullptr;
    if (esdguard::gmalloc(&p, 1000) != cudaSuccess) return ESD_ERR_CUDA;
    if (damage) cudaMemset(p + 1000, 0, 1);
    esdguard::gfree(p);
    return 1;
}

int esd_set_timing(esd_ctx* c, int32_t enable) {
    if (!c) return ESD_ERR_INVALID;
    c->timing = enable != 0;
    return ESD_OK;
}

int esd_kernel_time(esd_ctx* c, double* fused_ms, int64_t* fused_launches) {
    if (!c) return ESD_ERR_INVALID;
    int rc = esd_synchronize(c);
    if (rc) return rc;
    double total = 0.0;
    for (auto& ev : c->timing_events) {
        float ms = 0.f;
        CU(c, cudaEventElapsedTime(&ms, ev.first, ev.second));
        total += ms;
        cudaEventDestroy(ev.first);
        cudaEventDestroy(ev.second);
    }
    if (fused_ms) *fused_ms = total;
    if (fused_launches) *fused_launches = (int64_t)c->timing_events.size();
    c->timing_events.clear();
    return ESD_OK;
}

int64_t esd_kernel_launches(const esd_ctx* c) { return c ? c->launches : 0; }

}  // extern "C"
