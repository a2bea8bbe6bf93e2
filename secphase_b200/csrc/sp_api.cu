// sp_api.cu -- C ABI of libsecphase_b200.so (see include/secphase_b200.h) and the launcher
// that drives the kernels of sp_kernels.cuh.  Host side only plans, packs into pinned staging,
// enqueues, and turns the device results into the reference's tables; there is no CPU
// implementation of the hot path in this library.
#include <cuda.h>  // types of the driver's green-context API only; entry points come from cudaGetDriverEntryPoint
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/secphase_b200.h"
#include "sp_kernels.cuh"
#include "sp_plan.h"

static thread_local char g_err[512] = "";
static void set_err(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            set_err("CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__, __LINE__,  \
                    cudaGetErrorString(e_));                                                  \
            return SP_ECUDA;                                                                  \
        }                                                                                     \
    } while (0)

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return SP_OK;
        if (p) cudaFree(p);
        p = nullptr;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cap = 0;
            set_err("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
            return SP_ENOMEM;
        }
        cap = want;
        return SP_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};
struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return SP_OK;
        if (p) cudaFreeHost(p);
        p = nullptr;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e != cudaSuccess) {
            cap = 0;
            set_err("cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(e));
            return SP_ENOMEM;
        }
        cap = want;
        return SP_OK;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

enum { EV_START = 0, EV_H2D, EV_WALK, EV_GROUP, EV_EMIT, EV_HMM, EV_SCORE, EV_END, EV_N };

#define SP_N_AUX 9  // one per band class that a batch typically populates (3 slots x 10 streams stay under the 32 hardware queues)
struct Slot {
    cudaStream_t stream = nullptr;
    cudaStream_t aux[SP_N_AUX] = {};  // the HMM class launches fork onto these and join back
    cudaEvent_t ev_fork = nullptr, ev_join[SP_N_AUX] = {};
    cudaEvent_t ev[EV_N] = {};
    PinBuf h_in;    // staged batch: metadata always, pools only when the caller's are pageable
    DevBuf d_in;
    size_t in_bytes = 0, meta_bytes = 0, tag_pad_off = 0;
    struct Copy {
        const void *src;
        size_t off, bytes;
    } copies[4];
    int n_copies = 0;
    SpPlan plan;
    bool safe_caps = false;
    size_t off_gblk_cap = 0, off_gblk_off = 0, off_giv_off = 0;  // plan-derived sections inside h_in / d_in
    // device work tables
    DevBuf ops, imk, info, blk, iv, nb, gpos, ent, res, baq, gP, gout, gcnt, acnt, walk_fb, item_off, row_off, sdbl_off, score,
        fin_wide, fin, items, rows, order, bins, class_start, s_pool, fsave, gband, totals, work_counter, qual_out, set_base;
    // host results
    PinBuf h_tot, h_gout, h_score, h_info, h_fin, h_qual, h_rerun;
    const uint8_t *qual_zero_copy = nullptr;  // device view of the caller's page-locked quality pool, when it is read in place
    int64_t zero_copy_bytes = 0;
    bool fast_hmm = false;  // this batch's HMM launches used the fast kernel (+ strict re-run of flagged instances)
    size_t qual_bytes = 0;  // size of the batch's quality pool (== of qual_out in full_baq mode)
    bool full_baq = false;  // this batch was planned with SpConst::full_baq set
    SpBatchPtrs P;
    SpTotals tot;     // after the mid-pipeline read-back
    int state = 0;    // 0 idle, 1 uploaded, 2 in flight, 3 done
    // Two-phase enqueue: the caller's thread enqueues phase A (H2D, walk, group, scan, totals read-back)
    // and returns; the context's launcher thread waits for ev_a, sizes the HMM tables from the totals
    // and enqueues phase B (emit, sort, HMM, score, result copies).  phase: 0 none, 1 A enqueued,
    // 2 B enqueued (or failed: b_rc/b_err), guarded by sp_ctx::mu.
    cudaEvent_t ev_a = nullptr;
    cudaStream_t stream_hi = nullptr;  // phase B's integer kernels: same SMs as `stream`, scheduled ahead of the next batches' phase A
    cudaEvent_t ev_b = nullptr;        // end of phase B on stream_hi; `stream` waits for it
    int phase = 0;
    int b_rc = 0;
    std::string b_err;
    bool want_d2h = true;
    bool debug = false;
    int launches = 0;
    int64_t h2d_bytes = 0, d2h_bytes = 0;
    // host tables handed out by sp_wait / sp_debug_table
    std::vector<int32_t> r_group, r_extent, r_marker, dbg_rows;
    std::vector<int64_t> r_marker_off, dbg_off;
    std::vector<double> r_score;
};

#ifndef SP_WALK_WARP_MIN_OPS
#define SP_WALK_WARP_MIN_OPS 600  // planned op-table size from which an alignment gets a warp of its own (k_walk_warp)
#endif

struct sp_ctx {
    int device = 0;
    sp_params par;
    SpConst hC;
    DevBuf dC;
    DevBuf ref;          // codes
    DevBuf contig_off;   // int64[n+1]
    std::vector<int64_t> h_contig_off;
    int n_contigs = 0;
    Slot slot[SP_N_SLOTS];
    SpRng rng;
    bool debug_tables = false;
    int test_block_cap = 0;  // tests: first plan clamps every group's block workspace to this (forces the retry)
    int64_t cap_retries = 0;
    bool full_baq = false;  // sp_set_write_qual: --writeBam mode
    int group_mode = 0;  // SECPHASE_B200_GROUP: 0 default (lanes for groups of <= 4 alignments), 1 serial: thread per group, 2 lanes for all
    int score_mode = 0;  // SECPHASE_B200_SCORE: the same for K5
    bool walk_serial = false; // SECPHASE_B200_WALK=serial: thread-per-alignment walker only
    int walk_min_ops = SP_WALK_WARP_MIN_OPS;  // SECPHASE_B200_WALK=warp: 0 (every cs alignment gets a warp)
    bool hmm_merge = false; // SECPHASE_B200_HMM_MERGE=1: one fast launch for all classes (best pipelined, longest single-batch tail)
    int hmm_mode = 1;       // sp_set_hmm_mode: 0 strict (the reference's rounding order), 1 fast + guard band + strict re-run
    bool streams_ready = false;  // ensure_streams
    int sm_count = 0;
    size_t max_smem = 0;
    cudaEvent_t mark = nullptr;  // sp_mark / sp_elapsed_since_mark
    // SM partition (green contexts): the latency-bound integer kernels of a batch (walk, group, scan,
    // emit, sort, score) run on a few SMs of their own while the persistent FP64 HMM kernels of the
    // batches ahead of it own the rest; without it a persistent HMM CTA per SM keeps the 1024-thread
    // scan CTAs of the next batch waiting until the whole HMM kernel has drained.
    bool partitioned = false;
    CUgreenCtx g_int = nullptr, g_hmm = nullptr;
    int int_sms = 0, hmm_sms = 0;
    CUresult (*p_green_destroy)(CUgreenCtx) = nullptr;
    // launcher thread (phase B of every slot, in submission order)
    std::thread launcher;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::deque<int> queue;
    bool stop = false;
};

// ------------------------------------------------------------------------------------------
// SM partition.  Everything is looked up at run time so that the library neither links libcuda nor
// fails where green contexts are missing: any failure leaves the context un-partitioned.
static void sp_log(const char *fmt, ...) {
    if (!getenv("SECPHASE_B200_VERBOSE")) return;
    va_list ap;
    va_start(ap, fmt);
    fputs("[secphase_b200:info] ", stderr);
    vfprintf(stderr, fmt, ap);
    fputc('\n', stderr);
    va_end(ap);
}

template <class F> static bool drv(const char *name, F &fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
        cudaGetLastError();
        return false;
    }
    fn = reinterpret_cast<F>(p);
    return true;
}

static CUresult (*g_green_stream_create)(CUstream *, CUgreenCtx, unsigned int, int) = nullptr;

static void partition_sms(sp_ctx *c, int want_int) {
    c->partitioned = false;
    c->hmm_sms = c->sm_count;
    c->int_sms = c->sm_count;
    if (want_int <= 0) return;
    CUresult (*pDeviceGet)(CUdevice *, int) = nullptr;
    CUresult (*pGetRes)(CUdevice, CUdevResource *, CUdevResourceType) = nullptr;
    CUresult (*pSplit)(CUdevResource *, unsigned int *, const CUdevResource *, CUdevResource *, unsigned int, unsigned int) = nullptr;
    CUresult (*pDesc)(CUdevResourceDesc *, CUdevResource *, unsigned int) = nullptr;
    CUresult (*pCreate)(CUgreenCtx *, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
    if (!drv("cuDeviceGet", pDeviceGet) || !drv("cuDeviceGetDevResource", pGetRes) ||
        !drv("cuDevSmResourceSplitByCount", pSplit) || !drv("cuDevResourceGenerateDesc", pDesc) ||
        !drv("cuGreenCtxCreate", pCreate) || !drv("cuGreenCtxDestroy", c->p_green_destroy) ||
        !drv("cuGreenCtxStreamCreate", g_green_stream_create)) {
        sp_log("green-context entry points not available; SMs not partitioned");
        return;
    }
    CUdevice dev;
    CUdevResource all, small, rest;
    memset(&all, 0, sizeof(all)); memset(&small, 0, sizeof(small)); memset(&rest, 0, sizeof(rest));
    unsigned int nb = 1;
    CUresult r;
    if ((r = pDeviceGet(&dev, c->device)) != CUDA_SUCCESS || (r = pGetRes(dev, &all, CU_DEV_RESOURCE_TYPE_SM)) != CUDA_SUCCESS) {
        sp_log("cuDeviceGetDevResource failed (%d); SMs not partitioned", (int) r);
        return;
    }
    r = pSplit(&small, &nb, &all, &rest, CU_DEV_SM_RESOURCE_SPLIT_IGNORE_SM_COSCHEDULING, (unsigned) want_int);
    if (r != CUDA_SUCCESS || nb < 1 || small.sm.smCount == 0 || rest.sm.smCount == 0) {
        sp_log("cuDevSmResourceSplitByCount(%d) failed (%d); SMs not partitioned", want_int, (int) r);
        return;
    }
    CUdevResourceDesc d_int = nullptr, d_hmm = nullptr;
    if ((r = pDesc(&d_int, &small, 1)) != CUDA_SUCCESS || (r = pDesc(&d_hmm, &rest, 1)) != CUDA_SUCCESS) {
        sp_log("cuDevResourceGenerateDesc failed (%d); SMs not partitioned", (int) r);
        return;
    }
    if ((r = pCreate(&c->g_int, d_int, dev, CU_GREEN_CTX_DEFAULT_STREAM)) != CUDA_SUCCESS) {
        sp_log("cuGreenCtxCreate failed (%d); SMs not partitioned", (int) r);
        c->g_int = nullptr;
        return;
    }
    if ((r = pCreate(&c->g_hmm, d_hmm, dev, CU_GREEN_CTX_DEFAULT_STREAM)) != CUDA_SUCCESS) {
        sp_log("cuGreenCtxCreate failed (%d); SMs not partitioned", (int) r);
        c->p_green_destroy(c->g_int);
        c->g_int = c->g_hmm = nullptr;
        return;
    }
    c->partitioned = true;
    c->int_sms = (int) small.sm.smCount;
    c->hmm_sms = (int) rest.sm.smCount;
    sp_log("SM partition: %d SMs for the integer stages, %d SMs for the HMM kernels", c->int_sms, c->hmm_sms);
}

// a non-blocking stream inside green context g (or an ordinary one when the device is not partitioned)
static bool make_stream(sp_ctx *c, CUgreenCtx g, cudaStream_t *out, bool high_priority = false) {
    int prio = 0;
    if (high_priority) {
        int least = 0, greatest = 0;
        if (cudaDeviceGetStreamPriorityRange(&least, &greatest) == cudaSuccess) prio = greatest;
    }
    if (c->partitioned && g) {
        CUstream s = nullptr;
        if (g_green_stream_create(&s, g, CU_STREAM_NON_BLOCKING, prio) == CUDA_SUCCESS) {
            *out = reinterpret_cast<cudaStream_t>(s);
            return true;
        }
        return false;
    }
    return cudaStreamCreateWithPriority(out, cudaStreamNonBlocking, prio) == cudaSuccess;
}

// The SM partition and the streams inside it are created at the first batch, because the right split
// depends on the data: the integer stages (walk / group / emit) are thread-per-alignment latency
// chains whose cost grows with the cs/MD text, the HMM with the band cells.  HiFi read groups carry
// ~3 KB of tag text and are best served by 8 SMs of integer work against 140 of HMM; ONT groups carry
// ~17 KB and want 32 (measured: profiles/r01_sm_partition_sweep_v13.json; with the warp-per-alignment walker,
// whose time falls with the SMs it gets: profiles/r02_bench_sm_split_group_lanes_v30.txt).  SECPHASE_B200_INT_SMS=<n>
// overrides (0 = no partition); the driver rounds the request up to its own granularity.
static int ensure_streams(sp_ctx *c, const sp_flat_batch *hint) {
    if (c->streams_ready) return SP_OK;
    int want = 8;
    if (hint && hint->n_groups > 0 && hint->tag_off[hint->n_alns] / hint->n_groups > 8192) want = 32;
    if (const char *e = getenv("SECPHASE_B200_INT_SMS")) want = atoi(e);
    if (want >= c->sm_count) want = 0;
    partition_sms(c, want);
    bool ok = true;
    for (int s = 0; s < SP_N_SLOTS && ok; s++) {
        Slot &S = c->slot[s];
        ok = ok && make_stream(c, c->g_int, &S.stream);
        // With several batches in flight the integer SMs are kept busy by the walk/group kernels of the batches
        // behind; a batch's emit/sort (which gate its HMM launches) and score kernels must not queue behind them,
        // or the HMM SMs idle: they go to a stream of the highest priority.  SECPHASE_B200_NO_PRIORITY=1: one stream.
        if (!getenv("SECPHASE_B200_NO_PRIORITY")) ok = ok && make_stream(c, c->g_int, &S.stream_hi, true);
        for (int k = 0; k < SP_N_AUX && ok; k++) ok = ok && make_stream(c, c->g_hmm, &S.aux[k]);
    }
    if (!ok) {
        set_err("could not create streams: %s", cudaGetErrorString(cudaGetLastError()));
        return SP_ECUDA;
    }
    c->streams_ready = true;
    return SP_OK;
}

static void launcher_main(sp_ctx *c);

// ------------------------------------------------------------------------------------------
extern "C" {

const char *sp_last_error(void) { return g_err; }
const char *sp_version(void) { return "secphase_b200 0.1 (sm_100a)"; }

void sp_params_default(sp_params *p) {  // secphase.c:420-449
    p->baq_flag = 0;
    p->consensus = 0;
    p->indel_threshold = 10;
    p->min_q = 10;
    p->min_score = -10;
    p->set_q = 40;
    p->flank_margin = 500;
    p->prim_margin_score = 40;
    p->prim_margin_random = 0;
    p->conf_d = 1e-4;
    p->conf_e = 0.1;
    p->conf_b = 20;
}

int sp_params_preset(sp_params *p, const char *name) {  // secphase.c:477-504
    if (!p || !name) return SP_EINVAL;
    const bool hifi = strcmp(name, "hifi") == 0, ont = strcmp(name, "ont") == 0;
    if (!hifi && !ont) {
        set_err("unknown preset '%s'", name);
        return SP_EINVAL;
    }
    p->baq_flag = 1;
    p->consensus = 1;
    p->indel_threshold = hifi ? 10 : 20;
    p->conf_d = hifi ? 1e-4 : 1e-3;
    p->conf_e = 0.1;
    p->conf_b = 20;
    p->min_q = 10;
    p->set_q = hifi ? 40 : 20;
    p->prim_margin_score = hifi ? 40 : 20;
    p->prim_margin_random = 0;
    p->min_score = -10;
    return SP_OK;
}

sp_ctx *sp_create(const sp_params *p, int cuda_device) {
    if (!p) {
        set_err("sp_create: params is NULL");
        return nullptr;
    }
    // three slots x (1 + SP_N_AUX) streams: ask for enough hardware queues that streams of different
    // slots do not alias (read by the driver when it creates the device context, i.e. only if nothing
    // in the process has touched CUDA yet; bench.py and the CLI set it themselves at start-up)
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_err("no usable CUDA device (%s); libsecphase_b200 has no CPU path", cudaGetErrorString(e));
        return nullptr;
    }
    if (cuda_device < 0 || cuda_device >= ndev) {
        set_err("cuda_device %d out of range (0..%d)", cuda_device, ndev - 1);
        return nullptr;
    }
    if (cudaSetDevice(cuda_device) != cudaSuccess) {
        set_err("cudaSetDevice(%d) failed", cuda_device);
        return nullptr;
    }
    sp_ctx *c = new sp_ctx();
    c->device = cuda_device;
    c->par = *p;
    sp_fill_const(*p, c->hC);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, cuda_device);
    c->sm_count = prop.multiProcessorCount;
    c->max_smem = prop.sharedMemPerBlockOptin;
    if (c->dC.ensure(sizeof(SpConst)) != SP_OK ||
        cudaMemcpy(c->dC.p, &c->hC, sizeof(SpConst), cudaMemcpyHostToDevice) != cudaSuccess) {
        set_err("sp_create: could not upload constants");
        delete c;
        return nullptr;
    }
    bool attr_ok = cudaFuncSetAttribute(k_hmm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) c->max_smem) == cudaSuccess;
#define SP_ATTR(NW, NC)                                                                                                    \
    attr_ok = attr_ok &&                                                                                                   \
              cudaFuncSetAttribute(k_hmm2<NW, NC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) c->max_smem) == cudaSuccess && \
              cudaFuncSetAttribute(k_hmm2<NW, NC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) c->max_smem) == cudaSuccess
    SP_ATTR(1, 0); SP_ATTR(2, 0); SP_ATTR(3, 0);
    SP_ATTR(1, 41); SP_ATTR(1, 43); SP_ATTR(1, 45);
#if SP_N_EXACT_CLASSES > 3
    SP_ATTR(1, 47); SP_ATTR(1, 49); SP_ATTR(1, 51);
#endif
#undef SP_ATTR
#define SP_ATTRF(NC) attr_ok = attr_ok && cudaFuncSetAttribute(k_hmmf<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) c->max_smem) == cudaSuccess
    SP_ATTRF(41); SP_ATTRF(43); SP_ATTRF(45); SP_ATTRF(55); SP_ATTRF(73); SP_ATTRF(97); SP_ATTRF(125);
#undef SP_ATTRF
    if (const char *e = getenv("SECPHASE_B200_HMM")) c->hmm_mode = strcmp(e, "strict") == 0 ? 0 : 1;
    if (const char *e = getenv("SECPHASE_B200_GROUP")) c->group_mode = strcmp(e, "serial") == 0 ? 1 : strcmp(e, "lanes") == 0 ? 2 : 0;
    c->score_mode = c->group_mode;
    if (const char *e = getenv("SECPHASE_B200_SCORE")) c->score_mode = strcmp(e, "serial") == 0 ? 1 : strcmp(e, "lanes") == 0 ? 2 : 0;
    if (const char *e = getenv("SECPHASE_B200_WALK")) {
        c->walk_serial = strcmp(e, "serial") == 0;
        if (strcmp(e, "warp") == 0) c->walk_min_ops = 0;
        else if (atoi(e) > 0) c->walk_min_ops = atoi(e);
    }
    if (const char *e = getenv("SECPHASE_B200_HMM_MERGE")) c->hmm_merge = atoi(e) != 0;
    if (!attr_ok) {
        set_err("sp_create: cudaFuncSetAttribute(k_hmm) failed: %s", cudaGetErrorString(cudaGetLastError()));
        delete c;
        return nullptr;
    }
    bool ok = true;
    for (int s = 0; s < SP_N_SLOTS && ok; s++) {
        Slot &S = c->slot[s];
        for (int k = 0; k < EV_N; k++) ok = ok && cudaEventCreate(&S.ev[k]) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&S.ev_fork, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&S.ev_a, cudaEventDisableTiming | cudaEventBlockingSync) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&S.ev_b, cudaEventDisableTiming) == cudaSuccess;
        for (int k = 0; k < SP_N_AUX && ok; k++)
            ok = ok && cudaEventCreateWithFlags(&S.ev_join[k], cudaEventDisableTiming) == cudaSuccess;
    }
    if (!ok) {
        set_err("sp_create: could not create events: %s", cudaGetErrorString(cudaGetLastError()));
        sp_destroy(c);
        return nullptr;
    }
    c->rng.seed(1);
    c->launcher = std::thread(launcher_main, c);
    return c;
}

void sp_destroy(sp_ctx *c) {
    if (!c) return;
    if (c->launcher.joinable()) {
        {
            std::lock_guard<std::mutex> lk(c->mu);
            c->stop = true;
        }
        c->cv_work.notify_all();
        c->launcher.join();
    }
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
#ifdef SP_PROFILE_GROUP
    {
        unsigned long long h[16] = {};
        cudaMemcpyFromSymbol(h, sp_prof, sizeof(h));
        fprintf(stderr, "[secphase_b200:prof] thread-cycles: k_group markers %llu, consensus %llu, emit-count %llu; k_emit %llu\n",
                h[0], h[1], h[2], h[3]);
        fprintf(stderr, "[secphase_b200:prof] consensus rounds: sort+intersect blocks %llu, flank lists %llu, intersect flank %llu, "
                        "project %llu, sort %llu (+needs_to_find_blocks outside)\n", h[8], h[9], h[10], h[11], h[12]);
    }
#endif
    for (int s = 0; s < SP_N_SLOTS; s++) {
        Slot &S = c->slot[s];
        DevBuf *bufs[] = {&S.d_in, &S.ops, &S.imk, &S.info, &S.blk, &S.iv, &S.nb, &S.gpos, &S.ent, &S.res, &S.baq,
                          &S.gP, &S.gout, &S.gcnt, &S.acnt, &S.walk_fb, &S.item_off, &S.row_off, &S.sdbl_off, &S.score, &S.fin_wide,
                          &S.fin, &S.items, &S.rows, &S.order, &S.bins, &S.class_start, &S.s_pool, &S.fsave,
                          &S.gband, &S.totals, &S.work_counter, &S.qual_out, &S.set_base};
        for (DevBuf *b : bufs) b->release();
        PinBuf *pins[] = {&S.h_in, &S.h_tot, &S.h_gout, &S.h_score, &S.h_info, &S.h_fin, &S.h_qual, &S.h_rerun};
        for (PinBuf *b : pins) b->release();
        for (int k = 0; k < EV_N; k++)
            if (S.ev[k]) cudaEventDestroy(S.ev[k]);
        if (S.ev_fork) cudaEventDestroy(S.ev_fork);
        if (S.ev_a) cudaEventDestroy(S.ev_a);
        if (S.ev_b) cudaEventDestroy(S.ev_b);
        if (S.stream_hi) cudaStreamDestroy(S.stream_hi);
        for (int k = 0; k < SP_N_AUX; k++) {
            if (S.ev_join[k]) cudaEventDestroy(S.ev_join[k]);
            if (S.aux[k]) cudaStreamDestroy(S.aux[k]);
        }
        if (S.stream) cudaStreamDestroy(S.stream);
    }
    c->dC.release();
    c->ref.release();
    c->contig_off.release();
    if (c->mark) cudaEventDestroy(c->mark);
    if (c->partitioned && c->p_green_destroy) {
        c->p_green_destroy(c->g_int);
        c->p_green_destroy(c->g_hmm);
    }
    delete c;
}

void *sp_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
        set_err("sp_host_alloc(%zu): %s", bytes, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return p;
}
void sp_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

void sp_rng_seed(sp_ctx *c, unsigned seed) { c->rng.seed(seed); }
int sp_rng_next(sp_ctx *c) { return c->rng.next(); }

int sp_set_reference_codes(sp_ctx *c, int32_t n, const uint8_t *codes, const int64_t *off) {
    if (!c || n < 1 || !codes || !off) return SP_EINVAL;
    CK(cudaSetDevice(c->device));
    const int64_t total = off[n];
    int rc = c->ref.ensure((size_t) total + 64);
    if (rc) return rc;
    rc = c->contig_off.ensure(sizeof(int64_t) * (size_t) (n + 1));
    if (rc) return rc;
    CK(cudaMemcpy(c->ref.p, codes, (size_t) total, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->contig_off.p, off, sizeof(int64_t) * (size_t) (n + 1), cudaMemcpyHostToDevice));
    c->h_contig_off.assign(off, off + n + 1);
    c->n_contigs = n;
    return SP_OK;
}

int sp_set_reference_ascii(sp_ctx *c, int32_t n, const char *const *seqs, const int64_t *lens) {
    if (!c || n < 1 || !seqs || !lens) return SP_EINVAL;
    CK(cudaSetDevice(c->device));
    std::vector<int64_t> off((size_t) n + 1, 0);
    for (int i = 0; i < n; i++) off[(size_t) i + 1] = off[(size_t) i] + lens[i];
    const int64_t total = off[(size_t) n];
    int rc = c->ref.ensure((size_t) total + 64);
    if (rc) return rc;
    rc = c->contig_off.ensure(sizeof(int64_t) * (size_t) (n + 1));
    if (rc) return rc;
    // stage the ASCII text through a bounce buffer and encode on the device
    const size_t chunk = (size_t) 64 << 20;
    DevBuf tmp;
    rc = tmp.ensure(chunk);
    if (rc) return rc;
    for (int i = 0; i < n; i++) {
        for (int64_t o = 0; o < lens[i]; o += (int64_t) chunk) {
            const int64_t m = lens[i] - o < (int64_t) chunk ? lens[i] - o : (int64_t) chunk;
            CK(cudaMemcpy(tmp.p, seqs[i] + o, (size_t) m, cudaMemcpyHostToDevice));
            k_encode_ref<<<1024, 256>>>(tmp.as<uint8_t>(), c->ref.as<uint8_t>() + off[(size_t) i] + o, m);
            CK(cudaGetLastError());
            CK(cudaDeviceSynchronize());
        }
    }
    tmp.release();
    CK(cudaMemcpy(c->contig_off.p, off.data(), sizeof(int64_t) * (size_t) (n + 1), cudaMemcpyHostToDevice));
    c->h_contig_off = off;
    c->n_contigs = n;
    return SP_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------
// staging layout: every section 256-byte aligned inside one pinned buffer / one device buffer
struct Section {
    size_t off, bytes;
};
struct InLayout {
    Section grp_aln_off, flag, tid, pos, l_qseq, n_cigar, tag_kind, aln_grp, gblk_cap, glist;
    Section cigar_off, tag_off, seq_off, qual_off, ops_off, imk_off, gpos_off, gent_off, gblk_off, giv_off;
    Section cigar_pool, tag_pool, seq_pool, qual_pool;
    size_t total;
};

static bool is_pinned(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

static InLayout make_layout(const sp_flat_batch *b, const SpPlan &pl) {
    InLayout L;
    size_t o = 0;
    auto add = [&](Section &s, size_t bytes) {
        s.off = o;
        s.bytes = bytes;
        o = align_up(o + bytes + 16, 256);  // +16: SpByteReader may read one aligned 16-byte word past the end
    };
    const size_t G = (size_t) pl.G, A = (size_t) pl.A;
    add(L.grp_aln_off, 4 * (G + 1));
    add(L.flag, 4 * A); add(L.tid, 4 * A); add(L.pos, 4 * A); add(L.l_qseq, 4 * A); add(L.n_cigar, 4 * A);
    add(L.tag_kind, 4 * A); add(L.aln_grp, 4 * A); add(L.gblk_cap, 4 * G); add(L.glist, 4 * G);
    add(L.cigar_off, 8 * (A + 1)); add(L.tag_off, 8 * (A + 1)); add(L.seq_off, 8 * (A + 1)); add(L.qual_off, 8 * (A + 1));
    add(L.ops_off, 8 * (A + 1)); add(L.imk_off, 8 * (A + 1));
    add(L.gpos_off, 8 * (G + 1)); add(L.gent_off, 8 * (G + 1)); add(L.gblk_off, 8 * (G + 1)); add(L.giv_off, 8 * (G + 1));
    add(L.cigar_pool, 4 * (size_t) b->cigar_off[A]);
    add(L.tag_pool, (size_t) b->tag_off[A]);
    add(L.seq_pool, (size_t) b->seq_off[A]);
    add(L.qual_pool, (size_t) b->qual_off[A]);
    L.total = o;
    return L;
}

// zero_copy_ok: the caller's pools stay valid until sp_wait (sp_submit's contract), so a pool may be read in place
static int stage_batch(sp_ctx *c, Slot &S, const sp_flat_batch *b, bool zero_copy_ok = false) {
    // the plan scans every cs/MD byte once: a few threads when the batch carries a lot of tag text (ONT)
    int plan_threads = 1;
    if (b->n_alns > 0 && b->tag_off[b->n_alns] > ((int64_t) 4 << 20)) {
        const unsigned hw = std::thread::hardware_concurrency();
        plan_threads = hw >= 16 ? 8 : hw >= 4 ? (int) hw / 2 : 1;
    }
    int rc = sp_make_plan(b, c->par.indel_threshold, S.safe_caps, S.plan, plan_threads, c->test_block_cap);
    if (rc != SP_OK) {
        set_err("malformed batch (code %d): groups need 1..%d alignments, CIGAR ops limited to MIDSH=X", rc,
                SP_MAX_ALN_PER_GROUP);
        return rc;
    }
    const SpPlan &pl = S.plan;
    for (int a = 0; a < pl.A; a++) {
        if (b->tid[a] < 0 || b->tid[a] >= c->n_contigs) {
            set_err("alignment %d: tid %d outside the reference table (%d contigs)", a, b->tid[a], c->n_contigs);
            return SP_EINVAL;
        }
    }
    InLayout L = make_layout(b, pl);
    S.off_gblk_cap = L.gblk_cap.off;
    S.off_gblk_off = L.gblk_off.off;
    S.off_giv_off = L.giv_off.off;
    rc = S.h_in.ensure(L.total);
    if (rc) return rc;
    rc = S.d_in.ensure(L.total);
    if (rc) return rc;
    uint8_t *h = S.h_in.as<uint8_t>();
    const size_t G = (size_t) pl.G, A = (size_t) pl.A;
    auto put = [&](const Section &s, const void *src) { if (s.bytes) memcpy(h + s.off, src, s.bytes); };
    put(L.grp_aln_off, b->grp_aln_off);
    put(L.flag, b->flag); put(L.tid, b->tid); put(L.pos, b->pos); put(L.l_qseq, b->l_qseq); put(L.n_cigar, b->n_cigar);
    if (b->tag_kind) put(L.tag_kind, b->tag_kind); else memset(h + L.tag_kind.off, 0, L.tag_kind.bytes);
    put(L.aln_grp, pl.aln_grp.data()); put(L.gblk_cap, pl.gblk_cap.data()); put(L.glist, pl.glist.data());
    put(L.cigar_off, b->cigar_off); put(L.tag_off, b->tag_off); put(L.seq_off, b->seq_off); put(L.qual_off, b->qual_off);
    put(L.ops_off, pl.ops_off.data()); put(L.imk_off, pl.imk_off.data());
    put(L.gpos_off, pl.gpos_off.data()); put(L.gent_off, pl.gent_off.data());
    put(L.gblk_off, pl.gblk_off.data()); put(L.giv_off, pl.giv_off.data());
    // The four pools are the bulk of a batch.  A pool the caller keeps in page-locked memory
    // (sp_host_alloc, or any cudaHostAlloc/cudaHostRegister'ed range) is copied to the device
    // straight from where it lies; a pageable pool is first staged into the slot's pinned buffer.
    S.n_copies = 0;
    auto pool = [&](const Section &sec, const void *src) {
        if (sec.bytes == 0) return;
        Slot::Copy &cp = S.copies[S.n_copies++];
        cp.off = sec.off;
        cp.bytes = sec.bytes;
        if (is_pinned(src)) {
            cp.src = src;
        } else {
            memcpy(h + sec.off, src, sec.bytes);
            cp.src = h + sec.off;
        }
    };
    S.meta_bytes = L.cigar_pool.off;  // everything before the pools is small per-alignment metadata
    pool(L.cigar_pool, b->cigar_pool); pool(L.tag_pool, b->tag_pool); pool(L.seq_pool, b->seq_pool);
    // Raw qualities are two thirds of a HiFi batch, yet the marker path reads only the bases inside X runs (k_walk,
    // ptMarker.c:50-70) and one base per alignment at every marker position (k_group, ptMarker.c:91-105): a few dozen
    // bytes per alignment.  When the pool is page-locked (hence mapped into the device's address space) and the
    // batch is that sparse, the kernels read those bytes in place over PCIe instead of copying the whole pool.
    // Upper bound on the reads: per group (alignments) x (mismatched bases of all its alignments), known from the plan.
    S.qual_zero_copy = nullptr;
    S.zero_copy_bytes = 0;
    if (zero_copy_ok && !c->full_baq && L.qual_pool.bytes > 0 && !getenv("SECPHASE_B200_NO_ZERO_COPY") && is_pinned(b->qual_pool)) {
        int64_t reads = 0;
        for (int g = 0; g < pl.G; g++) reads += (int64_t) (b->grp_aln_off[g + 1] - b->grp_aln_off[g]) * pl.g_msum[(size_t) g];
        void *dp = nullptr;
        if (reads * 96 < (int64_t) L.qual_pool.bytes &&
            cudaHostGetDevicePointer(&dp, const_cast<uint8_t *>(b->qual_pool), 0) == cudaSuccess && dp) {
            S.qual_zero_copy = static_cast<const uint8_t *>(dp);
            S.zero_copy_bytes = reads * 32;  // one 32-byte sector per read
        } else {
            cudaGetLastError();
        }
    }
    if (!S.qual_zero_copy) pool(L.qual_pool, b->qual_pool);
    S.tag_pad_off = L.tag_pool.off + L.tag_pool.bytes;
    S.in_bytes = L.total;
    S.qual_bytes = L.qual_pool.bytes;
    S.full_baq = c->full_baq;
    if (S.full_baq) {
        if ((rc = S.qual_out.ensure(S.qual_bytes + 16))) return rc;
        if ((rc = S.h_qual.ensure(S.qual_bytes + 16))) return rc;
    }

    // device work tables
    const bool dbg = c->debug_tables;
    S.debug = dbg;
    if ((rc = S.ops.ensure(sizeof(SpOp) * (size_t) (pl.total_ops + 1)))) return rc;
    if ((rc = S.imk.ensure(sizeof(SpInitMarker) * (size_t) (pl.total_imk + 1)))) return rc;
    if ((rc = S.info.ensure(sizeof(SpAlnInfo) * (A + 1)))) return rc;
    if ((rc = S.blk.ensure(sizeof(SpBlock) * (size_t) (pl.total_blk + 1)))) return rc;
    if ((rc = S.iv.ensure(sizeof(SpIv) * (size_t) (pl.total_iv + 1)))) return rc;
    if ((rc = S.nb.ensure(4 * (A + 1)))) return rc;
    if ((rc = S.gpos.ensure(4 * (size_t) (pl.total_pos + 1)))) return rc;
    if ((rc = S.ent.ensure(sizeof(SpEntry) * (size_t) (pl.total_ent + 1)))) return rc;
    if ((rc = S.res.ensure(4 * (size_t) (pl.total_ent + 1)))) return rc;
    if (dbg && (rc = S.baq.ensure(4 * (size_t) (pl.total_ent + 1)))) return rc;
    if ((rc = S.gP.ensure(4 * (G + 1)))) return rc;
    if ((rc = S.gout.ensure(sizeof(SpGroupOut) * (G + 1)))) return rc;
    if ((rc = S.gcnt.ensure(sizeof(SpEmitCounts) * (G + 1)))) return rc;
    if ((rc = S.acnt.ensure(sizeof(SpEmitCounts) * (A + 1)))) return rc;
    if ((rc = S.walk_fb.ensure(4 * (A + 2)))) return rc;
    if ((rc = S.item_off.ensure(4 * (G + 2)))) return rc;
    if ((rc = S.row_off.ensure(4 * (G + 2)))) return rc;
    if ((rc = S.sdbl_off.ensure(8 * (G + 2)))) return rc;
    if ((rc = S.score.ensure(8 * (A + 1)))) return rc;
    if ((rc = S.fin_wide.ensure(24 * (size_t) (pl.total_ent + 1)))) return rc;
    if ((rc = S.fin.ensure(24 * (size_t) (pl.total_ent + 1)))) return rc;
    if ((rc = S.totals.ensure(sizeof(SpTotals)))) return rc;
    if ((rc = S.bins.ensure(4 * (size_t) (SP_SORT_BWBINS * SP_SORT_LBINS)))) return rc;
    if ((rc = S.class_start.ensure(4 * (SP_N_CLASSES + 3)))) return rc;
    if ((rc = S.h_tot.ensure(sizeof(SpTotals)))) return rc;
    if ((rc = S.h_gout.ensure(sizeof(SpGroupOut) * (G + 1)))) return rc;
    if ((rc = S.h_score.ensure(8 * (A + 1)))) return rc;
    if ((rc = S.h_info.ensure(sizeof(SpAlnInfo) * (A + 1)))) return rc;
    if ((rc = S.h_rerun.ensure(sizeof(int) * SP_N_CLASSES))) return rc;

    uint8_t *d = S.d_in.as<uint8_t>();
    SpBatchPtrs &P = S.P;
    memset(&P, 0, sizeof(P));
    P.G = pl.G;
    P.A = pl.A;
#define DP(T, sec) reinterpret_cast<const T *>(d + L.sec.off)
    P.grp_aln_off = DP(int32_t, grp_aln_off); P.flag = DP(int32_t, flag); P.tid = DP(int32_t, tid);
    P.pos = DP(int32_t, pos); P.l_qseq = DP(int32_t, l_qseq); P.n_cigar = DP(int32_t, n_cigar);
    P.tag_kind = DP(int32_t, tag_kind); P.aln_grp = DP(int32_t, aln_grp); P.gblk_cap = DP(int32_t, gblk_cap); P.glist = DP(int32_t, glist);
    P.cigar_off = DP(int64_t, cigar_off); P.tag_off = DP(int64_t, tag_off); P.seq_off = DP(int64_t, seq_off);
    P.qual_off = DP(int64_t, qual_off); P.ops_off = DP(int64_t, ops_off); P.imk_off = DP(int64_t, imk_off);
    P.gpos_off = DP(int64_t, gpos_off); P.gent_off = DP(int64_t, gent_off); P.gblk_off = DP(int64_t, gblk_off);
    P.giv_off = DP(int64_t, giv_off);
    P.cigar_pool = DP(uint32_t, cigar_pool); P.tag_pool = DP(uint8_t, tag_pool); P.seq_pool = DP(uint8_t, seq_pool);
    P.qual_pool = S.qual_zero_copy ? S.qual_zero_copy : DP(uint8_t, qual_pool);
#undef DP
    P.ops = S.ops.as<SpOp>(); P.imk = S.imk.as<SpInitMarker>(); P.info = S.info.as<SpAlnInfo>();
    P.blk = S.blk.as<SpBlock>(); P.iv = S.iv.as<SpIv>(); P.nb = S.nb.as<int32_t>(); P.gpos = S.gpos.as<int32_t>();
    P.ent = S.ent.as<SpEntry>(); P.res = S.res.as<int32_t>(); P.baq = dbg ? S.baq.as<int32_t>() : nullptr;
    P.gP = S.gP.as<int32_t>(); P.gout = S.gout.as<SpGroupOut>(); P.gcnt = S.gcnt.as<SpEmitCounts>(); P.acnt = S.acnt.as<SpEmitCounts>(); P.walk_fb = S.walk_fb.as<int32_t>();
    P.item_off = S.item_off.as<int32_t>(); P.row_off = S.row_off.as<int32_t>(); P.sdbl_off = S.sdbl_off.as<int64_t>();
    P.score = S.score.as<double>(); P.fin_wide = S.fin_wide.as<int32_t>(); P.fin = S.fin.as<int32_t>();
    P.ref = c->ref.as<uint8_t>(); P.contig_off = c->contig_off.as<int64_t>(); P.n_contigs = c->n_contigs;
    return SP_OK;
}

// K4 launches for instances already ordered by (band class, length).  Classes up to
// SP_H2_MAXBW run k_hmm2 (band in shared memory, sp_hmm2.cuh); wider bands run the generic k_hmm,
// whose band falls back to global memory when it outgrows shared memory.  Every class is an
// independent launch on its own auxiliary stream, widest (= fewest, slowest instances) first, so
// that a handful of wide instances overlaps with the bulk instead of trailing it.
template <int NW, int NC, bool IL>
static void launch_hmm2(sp_ctx *c, Slot &S, cudaStream_t st, int cls, int first, int cnt, const uint8_t *ref,
                        const uint8_t *qbytes, const uint8_t *seq_pool, const int64_t *seq_off, int64_t fs_stride,
                        const int64_t *set_base, int set_first, const int32_t *order = nullptr, const int *count_ptr = nullptr,
                        int *counter = nullptr) {
    // count_ptr != nullptr: strict re-run of the instances the fast kernel flagged -- their number is only known
    // on the device; cnt is then the upper bound (the class size) and a small grid is enough
    if (!order) order = S.order.as<int32_t>();
    if (!counter) counter = S.work_counter.as<int>() + cls;
    if (count_ptr && cnt > 32 * 64) cnt = 32 * 64;
    const int nblk = (cnt + 31) / 32;
    const int ncell = sp_h2_cells(sp_class_bw(cls));
    const size_t slab = (size_t) ncell * 32 * 24;
    int wpc = (int) (c->max_smem / slab);  // warps per CTA = band slabs per SM
    if (wpc > 7) wpc = 7;  // k_hmm2 is compiled for at most 224 threads per CTA
    if (wpc > nblk) wpc = nblk;
    int grid = (nblk + wpc - 1) / wpc;
    if (grid > c->hmm_sms) grid = c->hmm_sms;
    k_hmm2<NW, NC, IL><<<grid, 32 * wpc, slab * wpc, st>>>(c->dC.as<SpConst>(), S.items.as<SpItem>(), order,
                                                   first, cnt, ncell, ref, qbytes, seq_pool, seq_off,
                                                   S.s_pool.as<double>(), S.fsave.as<double>(),
                                                   set_base ? (int64_t) (2 * sp_class_bw(cls) + 1) * 64 : fs_stride,
                                                   S.rows.as<SpRow>(), counter, set_base, set_first, count_ptr);
}

template <int NC>
static void launch_hmmf(sp_ctx *c, Slot &S, cudaStream_t st, int cls, int first, int cnt, const uint8_t *ref,
                        const uint8_t *qbytes, const uint8_t *seq_pool, const int64_t *seq_off, int64_t fs_stride,
                        bool guard_all) {
    const int nblk = (cnt + 31) / 32;
    const size_t slab = (size_t) (NC + 2) * 32 * 16;
    int wpc = (int) (c->max_smem / slab);
    if (wpc > sp_hmmf_warps(NC)) wpc = sp_hmmf_warps(NC);  // k_hmmf<NC>'s launch bound
    // a small class spreads over as many SMs as it has sets (one warp each) instead of filling a few CTAs
    const int per_sm = (nblk + c->hmm_sms - 1) / c->hmm_sms;
    if (wpc > per_sm) wpc = per_sm;
    int grid = (nblk + wpc - 1) / wpc;
    if (grid > c->hmm_sms) grid = c->hmm_sms;
    k_hmmf<NC><<<grid, 32 * wpc, slab * wpc, st>>>(c->dC.as<SpConst>(), S.items.as<SpItem>(), S.order.as<int32_t>(), first, cnt,
                                                  ref, qbytes, seq_pool, seq_off, S.s_pool.as<double>(), S.fsave.as<double>(),
                                                  fs_stride, S.rows.as<SpRow>(), S.work_counter.as<int>() + cls,
                                                  S.work_counter.as<int>() + 2 * SP_N_CLASSES + cls, guard_all ? 1 : 0);
}

// `interleaved`: the -w mode's lane-interleaved forward-row pool (set bases from k_fs_sets), else
// every instance keeps its rows to itself at row0 * fs_stride
static int launch_hmm(sp_ctx *c, Slot &S, cudaStream_t st, const int32_t *class_count, int max_bw, const uint8_t *ref,
                      const uint8_t *qbytes, const uint8_t *seq_pool, const int64_t *seq_off, int64_t fs_stride,
                      bool interleaved = false, bool fast = false, bool guard_all = true) {
    int rc;
    // per class: work-queue head of the main launch, of the strict re-run launch, and the re-run count
    if ((rc = S.work_counter.ensure(sizeof(int) * 3 * SP_N_CLASSES))) return rc;
    CK(cudaMemsetAsync(S.work_counter.p, 0, sizeof(int) * 3 * SP_N_CLASSES, st));
    S.fast_hmm = fast;
    int first[SP_N_CLASSES + 1];
    first[0] = 0;
    for (int cls = 0; cls < SP_N_CLASSES; cls++) first[cls + 1] = first[cls] + class_count[cls];
    SpSetPlan sets;
    memset(&sets, 0, sizeof(sets));
    if (interleaved) {
        for (int cls = 0; cls < SP_N_CLASSES; cls++) {
            sets.first_item[cls] = first[cls];
            sets.count[cls] = class_count[cls];
            sets.cells[cls] = 2 * sp_class_bw(cls) + 1;
            sets.first_set[cls + 1] = sets.first_set[cls] + (class_count[cls] + 31) / 32;
        }
        if ((rc = S.set_base.ensure(8 * (size_t) (sets.first_set[SP_N_CLASSES] + 1)))) return rc;
        k_fs_sets<<<1, 1024, 0, st>>>(sets, S.items.as<SpItem>(), S.order.as<int32_t>(), S.set_base.as<int64_t>());
        S.launches++;
    }
    const int64_t *set_base = interleaved ? S.set_base.as<int64_t>() : nullptr;
    CK(cudaEventRecord(S.ev_fork, st));
    int used = 0;
    for (int cls = SP_N_CLASSES - 1; cls >= 0; cls--) {
        const int cnt = class_count[cls];
        if (cnt == 0) continue;
        cudaStream_t as = S.aux[used % SP_N_AUX];
        if (used < SP_N_AUX) CK(cudaStreamWaitEvent(as, S.ev_fork, 0));
        used++;
        const int bwc = sp_class_bw(cls);
        // the fast kernel wants sets of one width: always true for the exact classes, true enough for the bw <= 27
        // class, and for the wider ones only when the class is well populated (ONT) -- a sparsely filled wide class
        // (HiFi's few hundred instances over nine widths) keeps the strict kernel
        const int fcells = sp_hmmf_class_cells(cls);
        const bool fast_cls = fast && fcells != 0 && (fcells <= 64 || cnt >= 2048);
        if (fast_cls && (!c->hmm_merge || fcells > 64)) {
            switch (fcells) {
                case 41: launch_hmmf<41>(c, S, as, cls, first[cls], cnt, ref, qbytes, seq_pool, seq_off, fs_stride, guard_all); break;
                case 43: launch_hmmf<43>(c, S, as, cls, first[cls], cnt, ref, qbytes, seq_pool, seq_off, fs_stride, guard_all); break;
                case 45: launch_hmmf<45>(c, S, as, cls, first[cls], cnt, ref, qbytes, seq_pool, seq_off, fs_stride, guard_all); break;
                case 55: launch_hmmf<55>(c, S, as, cls, first[cls], cnt, ref, qbytes, seq_pool, seq_off, fs_stride, guard_all); break;
                case 73: launch_hmmf<73>(c, S, as, cls, first[cls], cnt, ref, qbytes, seq_pool, seq_off, fs_stride, guard_all); break;
                case 97: launch_hmmf<97>(c, S, as, cls, first[cls], cnt, ref, qbytes, seq_pool, seq_off, fs_stride, guard_all); break;
                default: launch_hmmf<125>(c, S, as, cls, first[cls], cnt, ref, qbytes, seq_pool, seq_off, fs_stride, guard_all); break;
            }
            S.launches++;
        } else if (fast_cls) {
            // SECPHASE_B200_HMM_MERGE=1: every class the fast kernel serves goes into ONE launch (they are a contiguous stretch of the order: the
            // band width is a run-time value of each warp, the slab is sized for the widest): one work queue, so no
            // class leaves SMs idle in a partial last wave of its own and only one tail remains.
            int lo = cls;
            while (lo > 0 && sp_hmmf_class_cells(lo - 1) != 0 && sp_hmmf_class_cells(lo - 1) <= 64) lo--;
            const int total = first[cls + 1] - first[lo];
            int widest = cls;
            while (widest > lo && class_count[widest] == 0) widest--;
            switch (sp_hmmf_class_cells(widest)) {
                case 41: launch_hmmf<41>(c, S, as, lo, first[lo], total, ref, qbytes, seq_pool, seq_off, fs_stride, guard_all); break;
                case 43: launch_hmmf<43>(c, S, as, lo, first[lo], total, ref, qbytes, seq_pool, seq_off, fs_stride, guard_all); break;
                case 45: launch_hmmf<45>(c, S, as, lo, first[lo], total, ref, qbytes, seq_pool, seq_off, fs_stride, guard_all); break;
                default: launch_hmmf<55>(c, S, as, lo, first[lo], total, ref, qbytes, seq_pool, seq_off, fs_stride, guard_all); break;
            }
            S.launches++;
            cls = lo;  // the classes below were part of this launch
        } else if (bwc != 0) {
            const int nw = sp_h2_words(bwc);
#define SP_LAUNCH(NW, NC)                                                                                          \
    do {                                                                                                           \
        if (interleaved)                                                                                           \
            launch_hmm2<NW, NC, true>(c, S, as, cls, first[cls], cnt, ref, qbytes, seq_pool, seq_off, fs_stride,  \
                                      set_base, sets.first_set[cls]);                                              \
        else                                                                                                       \
            launch_hmm2<NW, NC, false>(c, S, as, cls, first[cls], cnt, ref, qbytes, seq_pool, seq_off, fs_stride, \
                                       nullptr, 0);                                                                \
    } while (0)
            switch (sp_class_unrolled_cells(cls)) {
                case 41: SP_LAUNCH(1, 41); break;
                case 43: SP_LAUNCH(1, 43); break;
                case 45: SP_LAUNCH(1, 45); break;
#if SP_N_EXACT_CLASSES > 3
                case 47: SP_LAUNCH(1, 47); break;
                case 49: SP_LAUNCH(1, 49); break;
                case 51: SP_LAUNCH(1, 51); break;
#endif
                default:
                    if (nw == 1) SP_LAUNCH(1, 0);
                    else if (nw == 2) SP_LAUNCH(2, 0);
                    else SP_LAUNCH(3, 0);
            }
#undef SP_LAUNCH
        } else {
            const int nblk = (cnt + 31) / 32;
            int W = 2 * max_bw + 2;
            double *gband = nullptr;
            size_t smem = (size_t) W * 28 * 32;
            if (smem > c->max_smem) {
                if ((rc = S.gband.ensure((size_t) nblk * smem))) return rc;
                gband = S.gband.as<double>();
                smem = 0;
            }
            k_hmm<<<nblk, 32, smem, as>>>(c->dC.as<SpConst>(), S.items.as<SpItem>(), S.order.as<int32_t>(), first[cls],
                                          cnt, W, ref, qbytes, seq_pool, seq_off, S.s_pool.as<double>(),
                                          S.fsave.as<double>(), fs_stride, S.rows.as<SpRow>(), gband);
        }
        S.launches++;
    }
    for (int k = 0; k < used && k < SP_N_AUX; k++) {
        CK(cudaEventRecord(S.ev_join[k], S.aux[k]));
        CK(cudaStreamWaitEvent(st, S.ev_join[k], 0));
    }
    CK(cudaGetLastError());
    return SP_OK;
}

// Phase A, enqueued by the caller's thread for the batch staged in S (device copy already enqueued
// or resident): walk, group, scans, and the one mid-pipeline read-back (instance / row / band totals,
// which size the HMM tables and launches).  The launcher thread picks the slot up from there.
static int run_phase_a(sp_ctx *c, Slot &S) {
    cudaStream_t st = S.stream;
    const SpBatchPtrs &P = S.P;
    const SpConst *dC = c->dC.as<SpConst>();
    S.launches = 0;
    CK(cudaMemsetAsync(S.totals.p, 0, sizeof(SpTotals), st));
    CK(cudaMemsetAsync(S.res.p, 0xff, 4 * (size_t) (S.plan.total_ent + 1), st));  // SP_RES_RAW == -1
    if (P.A > 0) {
        if (c->walk_serial) {
            k_walk<<<(P.A + 127) / 128, 128, 0, st>>>(P, dC);
            S.launches++;
        } else {
            CK(cudaMemsetAsync(S.walk_fb.p, 0, 4, st));
            k_walk_warp<<<(P.A + 3) / 4, 128, 0, st>>>(P, dC, c->walk_min_ops);
            k_walk_list<<<(P.A + 127) / 128, 128, 0, st>>>(P, dC);
            S.launches += 2;
        }
    }
    CK(cudaEventRecord(S.ev[EV_WALK], st));
    if (P.G > 0) {
        // K2 + K3: a lane per alignment for read groups of up to four alignments (2 or 4 lanes a group); a thread per
        // group for the rest -- with 16 lanes a group the chains are shorter still, but the few SMs of the integer
        // partition then run out of issue slots (stress: 62 ms against 47 per 8192 groups on 8 SMs, profiles/
        // r02_bench_sm_split_group_lanes_v30.txt).  SECPHASE_B200_GROUP=serial|lanes forces one form for all.
        const int32_t *gn = S.plan.glist_n;
        const int lanes_to = c->group_mode == 1 ? 0 : c->group_mode == 2 ? 3 : 2;  // lane classes below this one get lanes
        int first = 0;
        for (int cls = 0; cls < 3; cls++) {
            const int cnt = gn[cls];
            if (cnt > 0) {
                if (cls < lanes_to) {
                    if (cls == 0) k_group_lanes<2><<<(cnt * 2 + 127) / 128, 128, 0, st>>>(P, dC, first, cnt);
                    else if (cls == 1) k_group_lanes<4><<<(cnt * 4 + 127) / 128, 128, 0, st>>>(P, dC, first, cnt);
                    else k_group_lanes<16><<<(cnt * 16 + 127) / 128, 128, 0, st>>>(P, dC, first, cnt);
                } else {
                    k_group<<<(cnt + 63) / 64, 64, 0, st>>>(P, dC, first, cnt);
                }
                S.launches++;
            }
            first += cnt;
        }
        k_count<<<(P.A + 127) / 128, 128, 0, st>>>(P, dC);
        k_group_counts<<<(P.G + 127) / 128, 128, 0, st>>>(P);
        S.launches += 2;
    }
    k_scan_groups<<<1, 1024, 0, st>>>(P, S.totals.as<SpTotals>());
    S.launches++;
    CK(cudaEventRecord(S.ev[EV_GROUP], st));
    // the one mid-pipeline read-back: instance / row / band totals size the HMM launch
    CK(cudaMemcpyAsync(S.h_tot.p, S.totals.p, sizeof(SpTotals), cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(S.ev_a, st));
    CK(cudaGetLastError());
    return SP_OK;
}

// Phase B, enqueued by the launcher thread once ev_a has fired: emit, sort, HMM, score.
static int run_phase_b(sp_ctx *c, Slot &S) {
    // (phase A is complete when this runs, so the high-priority stream needs no dependency on `stream`;
    // `stream` -- results, the next use of the slot -- waits for the end of phase B below)
    cudaStream_t st = S.stream_hi ? S.stream_hi : S.stream;
    const SpBatchPtrs &P = S.P;
    const SpConst *dC = c->dC.as<SpConst>();
    CK(cudaEventSynchronize(S.ev_a));
    S.tot = *S.h_tot.as<SpTotals>();
    const SpTotals &T = S.tot;
    int rc;
    if ((rc = S.items.ensure(sizeof(SpItem) * (size_t) (T.n_items + 1)))) return rc;
    if ((rc = S.rows.ensure(sizeof(SpRow) * (size_t) (T.n_rows + 1)))) return rc;
    if ((rc = S.order.ensure(4 * (size_t) (T.n_items + 1)))) return rc;
    if ((rc = S.s_pool.ensure(8 * (size_t) (T.s_doubles + 2)))) return rc;
    const bool fast = c->hmm_mode == 1 && !S.full_baq;
    // doubles per saved forward row: the widest band of the batch -- or, wider still, the cells of the fast
    // kernel's class that band falls in (its virtual band always has the class's full width)
    int64_t fs_cells = 2 * (int64_t) T.max_bw + 1;
    if (fast && sp_hmmf_class_cells(sp_band_class(T.max_bw)) > fs_cells) fs_cells = sp_hmmf_class_cells(sp_band_class(T.max_bw));
    const int64_t fs_stride = 2 * fs_cells;
    // -w mode: every lane of a warp saves and re-reads every row in lock step, so the forward rows of a
    // 32-instance set are interleaved lane by lane (coalesced 512-byte accesses) instead of each
    // instance keeping ~650-byte rows to itself.  Needs the shared-memory-band kernel for every
    // instance and window lengths the length sort resolves exactly (see k_fs_sets).
    const bool interleaved = S.full_baq && T.n_items > 0 && T.class_count[SP_N_CLASSES - 1] == 0 &&
                             T.max_lq <= SP_SORT_LBINS - 2 && !getenv("SECPHASE_B200_NO_INTERLEAVE");  // (tests: both layouts)
    if (interleaved) {
        int64_t dbl = 0;  // sum over sets of (longest instance's rows) <= max rows + class rows / 32, per class
        for (int cls = 0; cls + 1 < SP_N_CLASSES; cls++)
            if (T.class_count[cls] > 0)
                dbl += (int64_t) (2 * sp_class_bw(cls) + 1) * 64 * (T.max_lq + T.class_rows[cls] / 32 + 2);
        if ((rc = S.fsave.ensure(8 * (size_t) (dbl + 2)))) return rc;
    } else if ((rc = S.fsave.ensure(8 * (size_t) ((int64_t) T.n_rows * fs_stride + 2)))) {
        return rc;
    }
    if (P.G > 0 && T.n_items > 0) {
        k_emit<<<(P.A + 127) / 128, 128, 0, st>>>(P, dC, S.items.as<SpItem>(), S.rows.as<SpRow>());
        if (S.full_baq) {  // -w: k_emit laid out the windows only; their per-base rows are filled in parallel
            k_fill_rows<<<(T.n_items * 32 + 255) / 256, 256, 0, st>>>(P, S.items.as<SpItem>(), T.n_items, S.rows.as<SpRow>());
            S.launches++;
        }
        const int nbins = SP_SORT_BWBINS * SP_SORT_LBINS;
        CK(cudaMemsetAsync(S.bins.p, 0, 4 * (size_t) nbins, st));
        k_sort_hist<<<(T.n_items + 255) / 256, 256, 0, st>>>(S.items.as<SpItem>(), T.n_items, S.bins.as<int32_t>());
        k_sort_scan<<<1, 1024, 0, st>>>(S.bins.as<int32_t>(), nbins, S.class_start.as<int32_t>());
        k_sort_scatter<<<(T.n_items + 255) / 256, 256, 0, st>>>(S.items.as<SpItem>(), T.n_items, S.bins.as<int32_t>(),
                                                                  S.order.as<int32_t>());
        S.launches += 4;
    }
    CK(cudaEventRecord(S.ev[EV_EMIT], st));
    if (T.n_items > 0) {
        if ((rc = launch_hmm(c, S, st, T.class_count, T.max_bw, P.ref, nullptr, P.seq_pool, P.seq_off, fs_stride,
                             interleaved, fast, /*guard_all=*/false)))
            return rc;
    } else {
        S.fast_hmm = false;
    }
    CK(cudaEventRecord(S.ev[EV_HMM], st));
    if (S.full_baq && S.qual_bytes > 0) {
        // --writeBam: the records' quality arrays as calc_update_baq_all leaves them (ptMarker.c:786,
        // 709-720, 797-806): raw qualities, then every row of every HMM window, then the zeroed markers
        uint8_t *qo = S.qual_out.as<uint8_t>();
        CK(cudaMemcpyAsync(qo, P.qual_pool, S.qual_bytes, cudaMemcpyDeviceToDevice, st));
        if (T.n_rows > 0) {
            k_baq_rows<<<(T.n_rows + 255) / 256, 256, 0, st>>>(dC, S.items.as<SpItem>(), S.rows.as<SpRow>(), T.n_rows,
                                                             P.qual_off, P.qual_pool, qo);
            S.launches++;
        }
        if (P.G > 0 && T.n_items > 0) {
            k_baq_zero<<<(P.G + 63) / 64, 64, 0, st>>>(P, qo);
            S.launches++;
        }
    }
    if (P.G > 0) {
        const int32_t *gn = S.plan.glist_n;
        const double pm = c->par.prim_margin_score, ms = (double) c->par.min_score;
        // K5: a lane per alignment for every group size (16 lanes for groups of more than four alignments: 13.1 -> 4.6 ms
        // per 8192 stress groups on the 8 integer SMs); SECPHASE_B200_SCORE=serial: a thread per group
        const int lanes_to = c->score_mode == 1 ? 0 : 3;
        int first = 0;
        for (int cls = 0; cls < 3; cls++) {
            const int cnt = gn[cls];
            if (cnt > 0) {
                SpRow *rw = S.rows.as<SpRow>();
                SpTotals *tt = S.totals.as<SpTotals>();
                if (cls >= lanes_to) k_score<<<(cnt + 63) / 64, 64, 0, st>>>(P, dC, rw, pm, ms, tt, first, cnt);
                else if (cls == 0) k_score_lanes<2><<<(cnt * 2 + 127) / 128, 128, 0, st>>>(P, dC, rw, pm, ms, tt, first, cnt);
                else if (cls == 1) k_score_lanes<4><<<(cnt * 4 + 127) / 128, 128, 0, st>>>(P, dC, rw, pm, ms, tt, first, cnt);
                else k_score_lanes<16><<<(cnt * 16 + 127) / 128, 128, 0, st>>>(P, dC, rw, pm, ms, tt, first, cnt);
                S.launches++;
            }
            first += cnt;
        }
    }
    CK(cudaEventRecord(S.ev[EV_SCORE], st));
    if (st != S.stream) {
        CK(cudaEventRecord(S.ev_b, st));
        CK(cudaStreamWaitEvent(S.stream, S.ev_b, 0));
    }
    CK(cudaGetLastError());
    return SP_OK;
}

static int enqueue_results(Slot &S);

static void hand_to_launcher(sp_ctx *c, int slot) {
    {
        std::lock_guard<std::mutex> lk(c->mu);
        c->slot[slot].phase = 1;
        c->slot[slot].b_rc = 0;
        c->queue.push_back(slot);
    }
    c->cv_work.notify_one();
}

static void launcher_main(sp_ctx *c) {
    cudaSetDevice(c->device);
    for (;;) {
        int s;
        {
            std::unique_lock<std::mutex> lk(c->mu);
            c->cv_work.wait(lk, [&] { return c->stop || !c->queue.empty(); });
            if (c->queue.empty()) return;  // stop requested and nothing left to launch
            s = c->queue.front();
            c->queue.pop_front();
        }
        Slot &S = c->slot[s];
        int rc = run_phase_b(c, S);
        if (rc == SP_OK) rc = enqueue_results(S);
        {
            std::lock_guard<std::mutex> lk(c->mu);
            S.b_rc = rc;
            if (rc) S.b_err = g_err;  // this thread's message, re-raised by sp_wait in the caller's thread
            S.phase = 2;
        }
        c->cv_done.notify_all();
    }
}

// blocks until the launcher has enqueued (or failed to enqueue) phase B of the slot
static int await_phase_b(sp_ctx *c, Slot &S) {
    std::unique_lock<std::mutex> lk(c->mu);
    c->cv_done.wait(lk, [&] { return S.phase == 2; });
    if (S.b_rc) set_err("%s", S.b_err.c_str());
    return S.b_rc;
}

static int enqueue_h2d(Slot &S) {
    cudaStream_t st = S.stream;
    uint8_t *d = S.d_in.as<uint8_t>();
    CK(cudaMemcpyAsync(d, S.h_in.p, S.meta_bytes, cudaMemcpyHostToDevice, st));
    for (int k = 0; k < S.n_copies; k++)
        CK(cudaMemcpyAsync(d + S.copies[k].off, S.copies[k].src, S.copies[k].bytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(d + S.tag_pad_off, 0, 16, st));  // SpByteReader may read one aligned word past the tags
    return SP_OK;
}

static int enqueue_results(Slot &S) {
    cudaStream_t st = S.stream;
    const size_t G = (size_t) S.P.G, A = (size_t) S.P.A;
    CK(cudaMemcpyAsync(S.h_tot.p, S.totals.p, sizeof(SpTotals), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(S.h_gout.p, S.gout.p, sizeof(SpGroupOut) * G, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(S.h_score.p, S.score.p, 8 * A, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(S.h_info.p, S.info.p, sizeof(SpAlnInfo) * A, cudaMemcpyDeviceToHost, st));
    S.d2h_bytes = (int64_t) (sizeof(SpTotals) + sizeof(SpGroupOut) * G + 8 * A + sizeof(SpAlnInfo) * A);
    if (S.fast_hmm) {  // how many instances the fast kernel's guard band handed to the strict kernel, per class
        CK(cudaMemcpyAsync(S.h_rerun.p, S.work_counter.as<int>() + 2 * SP_N_CLASSES, sizeof(int) * SP_N_CLASSES,
                           cudaMemcpyDeviceToHost, st));
        S.d2h_bytes += (int64_t) sizeof(int) * SP_N_CLASSES;
    }
    if (S.full_baq && S.qual_bytes > 0) {
        CK(cudaMemcpyAsync(S.h_qual.p, S.qual_out.p, S.qual_bytes, cudaMemcpyDeviceToHost, st));
        S.d2h_bytes += (int64_t) S.qual_bytes;
    }
    return SP_OK;
}

// A batch whose kernels flagged SP_GERR_BLOCK_CAP (the tight per-group bound on consensus blocks was not
// enough, sp_plan.h) is laid out again with the provable bound and run once more from the walk.  The
// pools are already on the device and the per-group counts are kept in the plan, so this needs neither
// the caller's buffers nor another pass over the text: only the three plan-derived offset tables are
// rewritten and the metadata section is copied again.  Blocking; called from sp_wait.
static int rerun_with_safe_caps(sp_ctx *c, Slot &S) {
    S.safe_caps = true;
    uint8_t *h = S.h_in.as<uint8_t>();
    int rc = sp_plan_block_caps(S.plan, reinterpret_cast<const int32_t *>(h), true);  // grp_aln_off is the first section
    if (rc) {
        set_err("block workspace of the safe plan does not fit");
        return rc;
    }
    const SpPlan &pl = S.plan;
    const size_t G = (size_t) pl.G;
    memcpy(h + S.off_gblk_cap, pl.gblk_cap.data(), 4 * G);
    memcpy(h + S.off_gblk_off, pl.gblk_off.data(), 8 * (G + 1));
    memcpy(h + S.off_giv_off, pl.giv_off.data(), 8 * (G + 1));
    if ((rc = S.blk.ensure(sizeof(SpBlock) * (size_t) (pl.total_blk + 1)))) return rc;
    if ((rc = S.iv.ensure(sizeof(SpIv) * (size_t) (pl.total_iv + 1)))) return rc;
    S.P.blk = S.blk.as<SpBlock>();
    S.P.iv = S.iv.as<SpIv>();
    CK(cudaMemcpyAsync(S.d_in.p, S.h_in.p, S.meta_bytes, cudaMemcpyHostToDevice, S.stream));
    const int launches = S.launches;
    if ((rc = run_phase_a(c, S))) return rc;
    hand_to_launcher(c, (int) (&S - c->slot));
    if ((rc = await_phase_b(c, S))) {
        cudaStreamSynchronize(S.stream);
        return rc;
    }
    CK(cudaStreamSynchronize(S.stream));
    S.launches += launches;
    c->cap_retries++;
    return SP_OK;
}

extern "C" {

int sp_submit(sp_ctx *c, const sp_flat_batch *b, int slot) {
    if (!c || !b || slot < 0 || slot >= SP_N_SLOTS) return SP_EINVAL;
    if (c->n_contigs == 0) {
        set_err("sp_submit before sp_set_reference_*");
        return SP_ESTATE;
    }
    CK(cudaSetDevice(c->device));
    if (int erc = ensure_streams(c, b)) return erc;
    Slot &S = c->slot[slot];
    if (S.state == 2) {
        set_err("slot %d still in flight; call sp_wait first", slot);
        return SP_ESTATE;
    }
    S.safe_caps = false;
    int rc = stage_batch(c, S, b, /*zero_copy_ok=*/true);
    if (rc) return rc;
    CK(cudaEventRecord(S.ev[EV_START], S.stream));
    if ((rc = enqueue_h2d(S))) return rc;
    CK(cudaEventRecord(S.ev[EV_H2D], S.stream));
    // bytes that cross to the device: what is copied, plus (an upper bound on) the sectors read in place
    S.h2d_bytes = (int64_t) S.meta_bytes + 16;
    for (int k = 0; k < S.n_copies; k++) S.h2d_bytes += (int64_t) S.copies[k].bytes;
    S.h2d_bytes += S.zero_copy_bytes;
    rc = run_phase_a(c, S);
    if (rc) return rc;
    S.state = 2;
    S.want_d2h = true;
    hand_to_launcher(c, slot);
    return SP_OK;
}

int sp_upload(sp_ctx *c, const sp_flat_batch *b, int slot) {
    if (!c || !b || slot < 0 || slot >= SP_N_SLOTS) return SP_EINVAL;
    if (c->n_contigs == 0) return SP_ESTATE;
    CK(cudaSetDevice(c->device));
    if (int erc = ensure_streams(c, b)) return erc;
    Slot &S = c->slot[slot];
    S.safe_caps = false;
    int rc = stage_batch(c, S, b);
    if (rc) return rc;
    if ((rc = enqueue_h2d(S))) return rc;
    CK(cudaStreamSynchronize(S.stream));
    S.h2d_bytes = 0;
    S.state = 1;
    return SP_OK;
}

int sp_run_resident(sp_ctx *c, int slot) {
    if (!c || slot < 0 || slot >= SP_N_SLOTS) return SP_EINVAL;
    CK(cudaSetDevice(c->device));
    Slot &S = c->slot[slot];
    if (S.state == 0) {
        set_err("slot %d has no uploaded batch", slot);
        return SP_ESTATE;
    }
    CK(cudaEventRecord(S.ev[EV_START], S.stream));
    CK(cudaEventRecord(S.ev[EV_H2D], S.stream));
    if (S.state == 2) {
        set_err("slot %d still in flight; call sp_wait first", slot);
        return SP_ESTATE;
    }
    int rc = run_phase_a(c, S);
    if (rc) return rc;
    S.state = 2;
    hand_to_launcher(c, slot);
    return SP_OK;
}

int sp_set_write_qual(sp_ctx *c, int on) {
    if (!c) return SP_EINVAL;
    for (int s = 0; s < SP_N_SLOTS; s++)
        if (c->slot[s].state == 2) {
            set_err("sp_set_write_qual: slot %d still has a batch in flight", s);
            return SP_ESTATE;
        }
    CK(cudaSetDevice(c->device));
    c->full_baq = on != 0;
    c->hC.full_baq = on != 0;
    CK(cudaMemcpy(c->dC.p, &c->hC, sizeof(SpConst), cudaMemcpyHostToDevice));
    // staged / resident batches were planned (tables, row counts) for the other mode: they must be
    // uploaded again before they can run, and their result tables are no longer handed out
    for (int s = 0; s < SP_N_SLOTS; s++)
        if (c->slot[s].state == 1 || c->slot[s].state == 3) c->slot[s].state = 0;
    return SP_OK;
}

int sp_set_hmm_mode(sp_ctx *c, int mode) {
    if (!c || (mode != 0 && mode != 1)) return SP_EINVAL;
    for (int s = 0; s < SP_N_SLOTS; s++)
        if (c->slot[s].state == 2) {
            set_err("sp_set_hmm_mode: slot %d still has a batch in flight", s);
            return SP_ESTATE;
        }
    c->hmm_mode = mode;
    return SP_OK;
}
int sp_get_hmm_mode(sp_ctx *c) { return c ? c->hmm_mode : SP_EINVAL; }

int sp_mark(sp_ctx *c) {
    if (!c) return SP_EINVAL;
    CK(cudaSetDevice(c->device));
    if (int erc = ensure_streams(c, nullptr)) return erc;
    if (!c->mark) CK(cudaEventCreate(&c->mark));
    CK(cudaEventRecord(c->mark, c->slot[0].stream));
    return SP_OK;
}

int sp_elapsed_since_mark(sp_ctx *c, int slot, float *ms) {
    if (!c || slot < 0 || slot >= SP_N_SLOTS || !ms) return SP_EINVAL;
    if (!c->mark || c->slot[slot].state != 3) {
        set_err("sp_elapsed_since_mark needs sp_mark and a completed sp_wait on slot %d", slot);
        return SP_ESTATE;
    }
    CK(cudaSetDevice(c->device));
    CK(cudaEventElapsedTime(ms, c->mark, c->slot[slot].ev[EV_END]));
    return SP_OK;
}

int sp_sm_partition(sp_ctx *c, int32_t *int_sms, int32_t *hmm_sms) {
    if (!c) return SP_EINVAL;
    if (!c->streams_ready) {  // decided at the first batch; before that report the request a HiFi batch would get
        if (cudaSetDevice(c->device) != cudaSuccess) return SP_ECUDA;
        if (int erc = ensure_streams(c, nullptr)) return erc;
    }
    if (int_sms) *int_sms = c->int_sms;
    if (hmm_sms) *hmm_sms = c->hmm_sms;
    return c->partitioned ? 1 : 0;
}

int sp_poll(sp_ctx *c, int slot) {
    if (!c || slot < 0 || slot >= SP_N_SLOTS) return SP_EINVAL;
    Slot &S = c->slot[slot];
    if (S.state != 2) {
        set_err("slot %d has nothing in flight", slot);
        return SP_ESTATE;
    }
    {
        std::lock_guard<std::mutex> lk(c->mu);
        if (S.phase != 2) return 0;
        if (S.b_rc) return 1;  // sp_wait reports the error
    }
    CK(cudaSetDevice(c->device));
    cudaError_t e = cudaStreamQuery(S.stream);
    if (e == cudaSuccess) return 1;
    if (e == cudaErrorNotReady) {
        cudaGetLastError();
        return 0;
    }
    set_err("cudaStreamQuery: %s", cudaGetErrorString(e));
    return SP_ECUDA;
}

int sp_wait(sp_ctx *c, int slot, sp_result *out) {
    if (!c || slot < 0 || slot >= SP_N_SLOTS || !out) return SP_EINVAL;
    CK(cudaSetDevice(c->device));
    Slot &S = c->slot[slot];
    if (S.state != 2) {
        set_err("slot %d has nothing in flight", slot);
        return SP_ESTATE;
    }
    if (int brc = await_phase_b(c, S)) {
        cudaStreamSynchronize(S.stream);
        S.state = 3;
        return brc;
    }
    CK(cudaStreamSynchronize(S.stream));
    SpTotals T = *S.h_tot.as<SpTotals>();
    if ((T.err & SP_GERR_BLOCK_CAP) && !(T.err & SP_GERR_BADOP) && !S.safe_caps) {
        if (int rrc = rerun_with_safe_caps(c, S)) {
            S.state = 3;
            return rrc;
        }
        T = *S.h_tot.as<SpTotals>();
    }
    // second, exact-size copy: the compact final-marker table
    int rc = S.h_fin.ensure(24 * (size_t) (T.fin_rows + 1));
    if (rc) return rc;
    if (T.fin_rows > 0)
        CK(cudaMemcpyAsync(S.h_fin.p, S.fin.p, 24 * (size_t) T.fin_rows, cudaMemcpyDeviceToHost, S.stream));
    CK(cudaEventRecord(S.ev[EV_END], S.stream));
    CK(cudaStreamSynchronize(S.stream));
    S.d2h_bytes += 24 * (int64_t) T.fin_rows;
    S.state = 3;
    if (T.err & SP_GERR_BADOP) {
        set_err("batch contains CIGAR operations the reference does not define (N/P)");
        return SP_EUNSUPPORTED;
    }
    if (T.err & (SP_GERR_BLOCK_CAP | SP_GERR_MARKER_CAP | SP_GERR_OP_CAP)) {
        set_err("internal table bound exceeded (flags %d)", T.err);
        return SP_ECAPACITY;
    }
    const int G = S.P.G, A = S.P.A;
    const SpGroupOut *go = S.h_gout.as<SpGroupOut>();
    const double *sc = S.h_score.as<double>();
    const SpAlnInfo *inf = S.h_info.as<SpAlnInfo>();
    const int32_t *fin = S.h_fin.as<int32_t>();
    const int32_t *gao = S.h_in.as<int32_t>();  // grp_aln_off is the first staged section
    S.r_group.resize((size_t) G * SP_GROUP_W);
    S.r_marker_off.resize((size_t) G + 1);
    S.r_marker.resize((size_t) T.fin_rows * SP_MARKER_W);
    S.r_score.assign(sc, sc + A);
    S.r_extent.resize((size_t) A * 4);
    for (int a = 0; a < A; a++) {
        S.r_extent[(size_t) a * 4 + 0] = inf[a].rfs;
        S.r_extent[(size_t) a * 4 + 1] = inf[a].rfe;
        S.r_extent[(size_t) a * 4 + 2] = inf[a].rds_f;
        S.r_extent[(size_t) a * 4 + 3] = inf[a].rde_f;
    }
    int64_t mo = 0;
    for (int g = 0; g < G; g++) {
        const SpGroupOut &o = go[g];
        const int a0 = gao[g], n = gao[g + 1] - a0;
        // tie-break RNG replay in submission order (ptAlignment.c:156-171)
        const int best = sp_finalize_best(c->rng, n, sc + a0, o.prim_idx, o.max_idx, o.tie_mask,
                                          c->par.prim_margin_score, (double) c->par.min_score,
                                          c->par.prim_margin_random);
        int32_t *row = &S.r_group[(size_t) g * SP_GROUP_W];
        row[0] = best; row[1] = o.prim_idx; row[2] = o.n_init; row[3] = o.n_after_allmm; row[4] = o.n_filled;
        row[5] = o.n_after_ins; row[6] = o.margin_eff; row[7] = o.conf_len; row[8] = o.n_final; row[9] = o.scored;
        S.r_marker_off[(size_t) g] = mo;
        if (o.n_final > 0)
            memcpy(&S.r_marker[(size_t) mo * SP_MARKER_W], fin + (size_t) o.fin_off * 6, 24 * (size_t) o.n_final);
        mo += o.n_final;
    }
    S.r_marker_off[(size_t) G] = mo;
    memset(out, 0, sizeof(*out));
    out->n_groups = G;
    out->n_alns = A;
    out->group = S.r_group.data();
    out->score = S.r_score.data();
    out->extent = S.r_extent.data();
    out->marker_off = S.r_marker_off.data();
    out->marker = S.r_marker.data();
    out->hmm_instances = T.n_items;
    out->hmm_cells = T.cells;
    out->h2d_bytes = S.h2d_bytes;
    out->d2h_bytes = S.d2h_bytes;
    out->gpu_launches = S.launches;
    out->hmm_mode = S.fast_hmm ? 1 : 0;
    out->hmm_strict_reruns = 0;
    if (S.fast_hmm && T.n_items > 0)
        for (int k = 0; k < SP_N_CLASSES; k++) out->hmm_strict_reruns += S.h_rerun.as<int>()[k];
    out->baq_qual = S.full_baq ? S.h_qual.as<uint8_t>() : nullptr;
    out->baq_qual_bytes = S.full_baq ? (int64_t) S.qual_bytes : 0;
    float ms = 0;
    const int order[8] = {EV_START, EV_H2D, EV_WALK, EV_GROUP, EV_EMIT, EV_HMM, EV_SCORE, EV_END};
    for (int k = 0; k < 7; k++) {
        ms = 0;
        cudaEventElapsedTime(&ms, S.ev[order[k]], S.ev[order[k + 1]]);
        out->ms_stage[k] = ms;
    }
    out->ms_stage[7] = 0;
    cudaEventElapsedTime(&ms, S.ev[EV_START], S.ev[EV_END]);
    out->ms_total = ms;
    out->ms_hmm = out->ms_stage[4];
    return SP_OK;
}

int64_t sp_debug_table(sp_ctx *c, int slot, int what, const int32_t **rows_out, const int64_t **off_out) {
    if (!c || slot < 0 || slot >= SP_N_SLOTS) return SP_EINVAL;
    if (what == -1) {  // switch: keep the post-BAQ qualities of later batches
        c->debug_tables = true;
        return 0;
    }
    if (what <= -100) {  // tests: clamp the first plan's block workspaces to -(what+100) entries per list
        c->test_block_cap = -(what + 100);
        return 0;
    }
    if (what == -2) return c->cap_retries;  // batches re-run with the provable bounds so far
    if (what == -3) {  // alignments of the slot's last batch the warp walker left to the serial one (-1: serial mode)
        if (c->walk_serial) return -1;
        if (!c->slot[slot].walk_fb.p) return 0;
        int32_t n = 0;
        if (cudaSetDevice(c->device) != cudaSuccess ||
            cudaMemcpy(&n, c->slot[slot].walk_fb.p, 4, cudaMemcpyDeviceToHost) != cudaSuccess)
            return SP_ECUDA;
        return n;
    }
    if (!rows_out) return SP_EINVAL;
    if (cudaSetDevice(c->device) != cudaSuccess) return SP_ECUDA;
    Slot &S = c->slot[slot];
    if (S.state != 3) {
        set_err("sp_debug_table: slot %d has no completed batch", slot);
        return SP_ESTATE;
    }
    const SpPlan &pl = S.plan;
    const int G = pl.G, A = pl.A;
    const int32_t *gao = S.h_in.as<int32_t>();
    S.dbg_rows.clear();
    S.dbg_off.clear();
    if (off_out) *off_out = nullptr;
#define D2H(vec, T, buf, count)                                                                            \
    std::vector<T> vec((size_t) (count) + 1);                                                              \
    if ((count) > 0 && cudaMemcpy(vec.data(), (buf).p, sizeof(T) * (size_t) (count), cudaMemcpyDeviceToHost) != cudaSuccess) \
        return SP_ECUDA;
    if (what == 0 || what == 1) {
        if (what == 1 && !S.debug) {
            set_err("post-BAQ table needs sp_debug_table(ctx,0,-1,...) before the submit");
            return SP_ESTATE;
        }
        D2H(gP, int32_t, S.gP, G);
        D2H(gpos, int32_t, S.gpos, pl.total_pos);
        D2H(ent, SpEntry, S.ent, pl.total_ent);
        std::vector<int32_t> baq;
        if (what == 1) {
            baq.resize((size_t) pl.total_ent + 1);
            if (pl.total_ent > 0 &&
                cudaMemcpy(baq.data(), S.baq.p, 4 * (size_t) pl.total_ent, cudaMemcpyDeviceToHost) != cudaSuccess)
                return SP_ECUDA;
        }
        S.dbg_off.push_back(0);
        for (int g = 0; g < G; g++) {
            const int n = gao[g + 1] - gao[g];
            for (int p = 0; p < gP[(size_t) g]; p++)
                for (int i = 0; i < n; i++) {
                    const size_t e = (size_t) pl.gent_off[(size_t) g] + (size_t) p * n + i;
                    const int32_t row[6] = {i, gpos[(size_t) pl.gpos_off[(size_t) g] + p], ent[e].base_idx,
                                            what == 1 ? baq[e] : ent[e].q, ent[e].flags & 1, ent[e].ref_pos};
                    S.dbg_rows.insert(S.dbg_rows.end(), row, row + 6);
                }
            S.dbg_off.push_back((int64_t) S.dbg_rows.size() / 6);
        }
        *rows_out = S.dbg_rows.data();
        if (off_out) *off_out = S.dbg_off.data();
        return (int64_t) S.dbg_rows.size() / 6;
    }
    if (what == 2) {
        D2H(nb, int32_t, S.nb, A);
        D2H(blk, SpBlock, S.blk, pl.total_blk);
        S.dbg_off.push_back(0);
        for (int g = 0; g < G; g++) {
            const int a0 = gao[g], n = gao[g + 1] - a0;
            for (int i = 0; i < n; i++) {
                const SpBlock *bl = blk.data() + pl.gblk_off[(size_t) g] + (int64_t) i * pl.gblk_cap[(size_t) g];
                for (int k = 0; k < nb[(size_t) (a0 + i)]; k++) {
                    const int32_t row[6] = {bl[k].rfs, bl[k].rfe, bl[k].sqs, bl[k].sqe, bl[k].rds_f, bl[k].rde_f};
                    S.dbg_rows.insert(S.dbg_rows.end(), row, row + 6);
                }
                S.dbg_off.push_back((int64_t) S.dbg_rows.size() / 6);
            }
        }
        *rows_out = S.dbg_rows.data();
        if (off_out) *off_out = S.dbg_off.data();
        return (int64_t) S.dbg_rows.size() / 6;
    }
    if (what == 3) {
        D2H(items, SpItem, S.items, S.tot.n_items);
        for (int k = 0; k < S.tot.n_items; k++) {
            const SpItem &I = items[(size_t) k];
            const int32_t row[SP_HMM_W] = {I.aln, I.l_ref, I.l_query, I.par_bw, I.blk, I.row0, I.n_rows, 0};
            S.dbg_rows.insert(S.dbg_rows.end(), row, row + SP_HMM_W);
        }
        *rows_out = S.dbg_rows.data();
        return S.tot.n_items;
    }
    if (what == 4) {
        D2H(rows, SpRow, S.rows, S.tot.n_rows);
        for (int k = 0; k < S.tot.n_rows; k++) {
            const int32_t row[4] = {rows[(size_t) k].item, rows[(size_t) k].t, rows[(size_t) k].state, rows[(size_t) k].q};
            S.dbg_rows.insert(S.dbg_rows.end(), row, row + 4);
        }
        *rows_out = S.dbg_rows.data();
        return S.tot.n_rows;
    }
#undef D2H
    return SP_EINVAL;
}

int sp_hmm_batch(sp_ctx *c, int32_t n, const uint8_t *ref_pool, const int64_t *ref_off, const int32_t *l_ref,
                 const uint8_t *query_pool, const int64_t *query_off, const int32_t *l_query, const int32_t *par_bw,
                 const int64_t *row_off, const int32_t *rows_t, int32_t *state_out, uint8_t *q_out, double *pmax_out,
                 float *ms_kernel) {
    if (!c || n < 0 || !ref_pool || !ref_off || !l_ref || !query_pool || !query_off || !l_query || !par_bw ||
        !row_off || !rows_t || !state_out || !q_out)
        return SP_EINVAL;
    if (n == 0) return SP_OK;
    CK(cudaSetDevice(c->device));
    // instance + row tables
    std::vector<SpItem> items((size_t) n);
    const int64_t n_rows = row_off[n];
    std::vector<SpRow> rows((size_t) n_rows + 1);
    int64_t ref_total = 0, q_total = 0, s_total = 0;
    int max_bw = 0;
    int cls_count[SP_N_CLASSES];
    for (int k = 0; k < SP_N_CLASSES; k++) cls_count[k] = 0;
    for (int j = 0; j < n; j++) {
        if (l_ref[j] <= 0 || l_query[j] <= 0) {
            set_err("sp_hmm_batch: instance %d has an empty sequence", j);
            return SP_EINVAL;
        }
        SpItem &I = items[(size_t) j];
        I.ref_off = ref_off[j];
        I.aln = -1;
        I.blk = 0;
        I.l_ref = l_ref[j];
        I.l_query = l_query[j];
        I.q_sqs = 0;
        I.par_bw = par_bw[j];
        I.row0 = (int32_t) row_off[j];
        I.n_rows = (int32_t) (row_off[j + 1] - row_off[j]);
        I.query_off = query_off[j];
        I.op_first = 0;
        I.op_last = -1;
        I.s_off = s_total;
        s_total += l_query[j] + 2;
        if (ref_off[j] + l_ref[j] > ref_total) ref_total = ref_off[j] + l_ref[j];
        if (query_off[j] + l_query[j] > q_total) q_total = query_off[j] + l_query[j];
        const int bw = sp_hmm_bw(I.l_ref, I.l_query, I.par_bw);
        if (bw > max_bw) max_bw = bw;
        if (I.n_rows > 0) cls_count[sp_band_class(bw)]++;
        for (int64_t r = row_off[j]; r < row_off[j + 1]; r++) {
            SpRow &R = rows[(size_t) r];
            R.item = j;
            R.t = rows_t[r];
            R.entry = -1;
            R.expected = 0;
            R.state = 0;
            R.q = 0;
            R.pmax = 0;
            if (R.t < 0 || R.t >= I.l_query || (r > row_off[j] && rows_t[r] <= rows_t[r - 1])) {
                set_err("sp_hmm_batch: rows of instance %d must be ascending and inside [0,l_query)", j);
                return SP_EINVAL;
            }
        }
    }
    if (int erc = ensure_streams(c, nullptr)) return erc;
    Slot &S = c->slot[0];
    if (S.state == 2) {
        set_err("sp_hmm_batch uses slot 0, which still has a batch in flight");
        return SP_ESTATE;
    }
    cudaStream_t st = S.stream;
    DevBuf d_ref, d_q;
    int rc;
    if ((rc = d_ref.ensure((size_t) ref_total + 16))) return rc;
    if ((rc = d_q.ensure((size_t) q_total + 16))) return rc;
    if ((rc = S.items.ensure(sizeof(SpItem) * (size_t) n))) return rc;
    if ((rc = S.rows.ensure(sizeof(SpRow) * (size_t) (n_rows + 1)))) return rc;
    if ((rc = S.order.ensure(4 * (size_t) n))) return rc;
    if ((rc = S.s_pool.ensure(8 * (size_t) (s_total + 2)))) return rc;
    const bool fast = c->hmm_mode == 1;
    int64_t fs_cells = 2 * (int64_t) max_bw + 1;
    if (fast && sp_hmmf_class_cells(sp_band_class(max_bw)) > fs_cells) fs_cells = sp_hmmf_class_cells(sp_band_class(max_bw));
    const int64_t fs_stride = 2 * fs_cells;
    if ((rc = S.fsave.ensure(8 * (size_t) (n_rows * fs_stride + 2)))) return rc;
    if ((rc = S.bins.ensure(4 * (size_t) (SP_SORT_BWBINS * SP_SORT_LBINS)))) return rc;
    if ((rc = S.class_start.ensure(4 * (SP_N_CLASSES + 3)))) return rc;
    CK(cudaMemcpyAsync(d_ref.p, ref_pool, (size_t) ref_total, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_q.p, query_pool, (size_t) q_total, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(S.items.p, items.data(), sizeof(SpItem) * (size_t) n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(S.rows.p, rows.data(), sizeof(SpRow) * (size_t) n_rows, cudaMemcpyHostToDevice, st));
    const int nbins = SP_SORT_BWBINS * SP_SORT_LBINS;
    CK(cudaMemsetAsync(S.bins.p, 0, 4 * (size_t) nbins, st));
    k_sort_hist<<<(n + 255) / 256, 256, 0, st>>>(S.items.as<SpItem>(), n, S.bins.as<int32_t>());
    k_sort_scan<<<1, 1024, 0, st>>>(S.bins.as<int32_t>(), nbins, S.class_start.as<int32_t>());
    k_sort_scatter<<<(n + 255) / 256, 256, 0, st>>>(S.items.as<SpItem>(), n, S.bins.as<int32_t>(), S.order.as<int32_t>());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    CK(cudaEventRecord(e0, st));
    const int launches_before = S.launches;
    if ((rc = launch_hmm(c, S, st, cls_count, max_bw, d_ref.as<uint8_t>(), d_q.as<uint8_t>(), nullptr, nullptr,
                         fs_stride, false, fast, /*guard_all=*/true)))
        return rc;
    S.launches = launches_before;
    CK(cudaEventRecord(e1, st));
    CK(cudaMemcpyAsync(rows.data(), S.rows.p, sizeof(SpRow) * (size_t) n_rows, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    if (ms_kernel) cudaEventElapsedTime(ms_kernel, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    for (int64_t r = 0; r < n_rows; r++) {
        state_out[r] = rows[(size_t) r].state;
        q_out[r] = (uint8_t) rows[(size_t) r].q;
        if (pmax_out) pmax_out[r] = rows[(size_t) r].pmax;
    }
    d_ref.release();
    d_q.release();
    return SP_OK;
}

int sp_fp64_peak(sp_ctx *c, int mode, double *ops_per_s, float *ms_out) {
    if (!c || !ops_per_s) return SP_EINVAL;
    CK(cudaSetDevice(c->device));
    const int blocks = c->sm_count * 8, threads = 256, iters = 1 << 16;
    DevBuf out;
    int rc = out.ensure(8 * (size_t) blocks * threads);
    if (rc) return rc;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_fp64_peak<<<blocks, threads>>>(out.as<double>(), 1024, mode);  // warm-up
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    k_fp64_peak<<<blocks, threads>>>(out.as<double>(), iters, mode);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    out.release();
    *ops_per_s = (double) blocks * threads * 8.0 * iters / (ms * 1e-3);
    if (ms_out) *ms_out = ms;
    return SP_OK;
}

}  // extern "C"
