// group.cu -- several GPUs behind the C ABI (include/treat_b200.h, "Several GPUs"): the device group of one process
// (one context + one host worker thread per GPU, contiguous read ranges, ncclCommInitAll) and the per-rank
// communicator of the one-process-per-GPU form.  The only collective of the path is an ncclAllReduce(uint64, sum) of the
// summary counters / heat map; the reads never cross GPUs.  NCCL is resolved with dlopen at first use.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "device_format.h"

using namespace treat_b200;

namespace {

// ---- the few NCCL entry points used, declared here so that neither nccl.h nor libnccl is needed to build ----------
typedef struct ncclComm *ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
static_assert(sizeof(ncclUniqueId) == TREAT_B200_COMM_ID_BYTES, "NCCL_UNIQUE_ID_BYTES");
constexpr int kNcclSuccess = 0, kNcclSum = 0, kNcclUint64 = 5;

struct Nccl {
    void *h = nullptr;
    int (*GetVersion)(int *) = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string err;
    bool ok = false;
};

Nccl &nccl()
{
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        // a copy already in the process (e.g. the one a Python host's torch brought) is reused; otherwise the system's
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            n.h = dlopen(name, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
            if (!n.h) n.h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (n.h) break;
        }
        if (!n.h) {
            const char *e = dlerror();
            n.err = std::string("libnccl.so.2 not found: ") + (e ? e : "");
            return;
        }
        bool all = true;
        auto sym = [&](const char *s) {
            void *p = dlsym(n.h, s);
            if (!p) {
                all = false;
                n.err = std::string("libnccl: missing symbol ") + s;
            }
            return p;
        };
        n.GetVersion = reinterpret_cast<decltype(n.GetVersion)>(sym("ncclGetVersion"));
        n.GetUniqueId = reinterpret_cast<decltype(n.GetUniqueId)>(sym("ncclGetUniqueId"));
        n.CommInitRank = reinterpret_cast<decltype(n.CommInitRank)>(sym("ncclCommInitRank"));
        n.CommInitAll = reinterpret_cast<decltype(n.CommInitAll)>(sym("ncclCommInitAll"));
        n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(sym("ncclCommDestroy"));
        n.AllReduce = reinterpret_cast<decltype(n.AllReduce)>(sym("ncclAllReduce"));
        n.GroupStart = reinterpret_cast<decltype(n.GroupStart)>(sym("ncclGroupStart"));
        n.GroupEnd = reinterpret_cast<decltype(n.GroupEnd)>(sym("ncclGroupEnd"));
        n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(sym("ncclGetErrorString"));
        n.ok = all;
    });
    return n;
}

std::string nccl_error(const char *what, int rc)
{
    Nccl &n = nccl();
    return std::string(what) + ": " + (n.GetErrorString ? n.GetErrorString(rc) : "NCCL error");
}

// one host thread per device: runs the jobs it is handed, one at a time
struct Worker {
    std::thread th;
    std::mutex m;
    std::condition_variable cv, done;
    std::function<int()> job;
    bool busy = false, stop = false;
    int result = 0;
    void start()
    {
        th = std::thread([this] {
            for (;;) {
                std::function<int()> j;
                {
                    std::unique_lock<std::mutex> lk(m);
                    cv.wait(lk, [&] { return stop || busy; });
                    if (stop) return;
                    j = job;
                }
                const int r = j();
                {
                    std::lock_guard<std::mutex> lk(m);
                    result = r;
                    busy = false;
                }
                done.notify_all();
            }
        });
    }
    void submit(std::function<int()> j)
    {
        {
            std::lock_guard<std::mutex> lk(m);
            job = std::move(j);
            busy = true;
        }
        cv.notify_all();
    }
    int wait()
    {
        std::unique_lock<std::mutex> lk(m);
        done.wait(lk, [&] { return !busy; });
        return result;
    }
    void shut()
    {
        {
            std::lock_guard<std::mutex> lk(m);
            stop = true;
        }
        cv.notify_all();
        if (th.joinable()) th.join();
    }
};

}  // namespace

struct treat_b200_group {
    int n = 0;
    std::vector<int> dev;
    std::vector<treat_b200_ctx *> ctx;
    std::vector<Worker *> worker;
    std::vector<ncclComm_t> comm;       // n > 1
    std::vector<cudaStream_t> stream;   // the collectives' stream of each device
    std::vector<uint64_t *> d_red;      // reduced summary / heat map of each device
    std::vector<size_t> red_cap;
    bool template_set = false;
    std::string err;
};

struct treat_b200_comm {
    treat_b200_ctx *ctx = nullptr;
    int device = 0, nranks = 1, rank = 0;
    ncclComm_t comm = nullptr;
    cudaStream_t stream = nullptr;  // the collectives run here, beside the caller's stream
    uint64_t *d_red = nullptr;
    size_t red_cap = 0;
    // treat_b200_comm_reduce_summary_device: two snapshots of the counters in turn, so that a reduction overlaps the next step
    uint64_t *d_snap[2] = {nullptr, nullptr};
    size_t snap_cap[2] = {0, 0};
    cudaEvent_t snap_done[2] = {nullptr, nullptr}, red_done[2] = {nullptr, nullptr};
    bool red_pending[2] = {false, false};
    unsigned turn = 0;
    std::string err;
};

namespace {

int gfail(treat_b200_group *g, int code, const std::string &msg)
{
    if (g) g->err = msg;
    return code;
}

// run fn(i) on every device's worker thread, return the first error (with that context's message)
int on_all(treat_b200_group *g, const std::function<int(int)> &fn)
{
    for (int i = 0; i < g->n; i++) g->worker[i]->submit([&fn, i] { return fn(i); });
    int rc = TREAT_B200_OK;
    for (int i = 0; i < g->n; i++) {
        const int r = g->worker[i]->wait();
        if (r != TREAT_B200_OK && rc == TREAT_B200_OK) {
            rc = r;
            g->err = "device " + std::to_string(g->dev[i]) + ": " + treat_b200_last_error(g->ctx[i]);
        }
    }
    return rc;
}

int reserve_red(uint64_t *&p, size_t &cap, size_t bytes)
{
    if (bytes <= cap) return TREAT_B200_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    if (cudaMalloc((void **)&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        return TREAT_B200_ERR_NOMEM;
    }
    cap = bytes;
    return TREAT_B200_OK;
}

// grouped all-reduce of one uint64 array per device (src[i] on device i), result read from device `from`
int group_allreduce(treat_b200_group *g, const std::vector<uint64_t *> &src, uint64_t len, int from, uint64_t *out)
{
    if (from < 0 || from >= g->n) return gfail(g, TREAT_B200_ERR_INVALID, "group: from_rank out of range");
    for (int i = 0; i < g->n; i++) {
        int rc = treat_b200_synchronize(g->ctx[i]);  // the counters are final
        if (rc) return gfail(g, rc, treat_b200_last_error(g->ctx[i]));
    }
    if (g->n == 1) {
        if (cudaSetDevice(g->dev[0]) != cudaSuccess || cudaMemcpy(out, src[0], len * 8, cudaMemcpyDeviceToHost) != cudaSuccess)
            return gfail(g, TREAT_B200_ERR_CUDA, std::string("group summary: ") + cudaGetErrorString(cudaGetLastError()));
        return TREAT_B200_OK;
    }
    Nccl &nc = nccl();
    for (int i = 0; i < g->n; i++) {
        if (cudaSetDevice(g->dev[i]) != cudaSuccess) return gfail(g, TREAT_B200_ERR_CUDA, "group: cudaSetDevice failed");
        int rc = reserve_red(g->d_red[i], g->red_cap[i], len * 8);
        if (rc) return gfail(g, rc, "group: out of device memory for the reduced counters");
    }
    int e = nc.GroupStart();
    for (int i = 0; i < g->n && e == kNcclSuccess; i++) {
        cudaSetDevice(g->dev[i]);
        e = nc.AllReduce(src[i], g->d_red[i], (size_t)len, kNcclUint64, kNcclSum, g->comm[i], g->stream[i]);
    }
    const int e2 = nc.GroupEnd();
    if (e != kNcclSuccess || e2 != kNcclSuccess) return gfail(g, TREAT_B200_ERR_CUDA, nccl_error("ncclAllReduce", e != kNcclSuccess ? e : e2));
    for (int i = 0; i < g->n; i++) {
        cudaSetDevice(g->dev[i]);
        if (cudaStreamSynchronize(g->stream[i]) != cudaSuccess)
            return gfail(g, TREAT_B200_ERR_CUDA, std::string("group all-reduce: ") + cudaGetErrorString(cudaGetLastError()));
    }
    cudaSetDevice(g->dev[from]);
    if (cudaMemcpy(out, g->d_red[from], len * 8, cudaMemcpyDeviceToHost) != cudaSuccess)
        return gfail(g, TREAT_B200_ERR_CUDA, std::string("group summary copy: ") + cudaGetErrorString(cudaGetLastError()));
    return TREAT_B200_OK;
}

}  // namespace

extern "C" {

int32_t treat_b200_nccl_version(void)
{
    Nccl &n = nccl();
    int v = 0;
    if (!n.ok || n.GetVersion(&v) != kNcclSuccess) return 0;
    return v;
}

int treat_b200_group_create(const int *devices, int32_t n, treat_b200_group **out)
{
    if (!out || n < 1 || n > 64) return TREAT_B200_ERR_INVALID;
    *out = nullptr;
    treat_b200_group *g = new treat_b200_group();
    g->n = n;
    for (int i = 0; i < n; i++) g->dev.push_back(devices ? devices[i] : i);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < i; j++)
            if (g->dev[i] == g->dev[j]) {
                delete g;
                return TREAT_B200_ERR_INVALID;  // one context per device
            }
    int rc = TREAT_B200_OK;
    for (int i = 0; i < n && rc == TREAT_B200_OK; i++) {
        treat_b200_ctx *c = nullptr;
        rc = treat_b200_create(g->dev[i], &c);
        if (rc == TREAT_B200_OK) g->ctx.push_back(c);
    }
    if (rc == TREAT_B200_OK && n > 1) {
        Nccl &nc = nccl();
        if (!nc.ok) {
            fprintf(stderr, "[treat_b200] %s\n", nc.err.c_str());
            rc = TREAT_B200_ERR_NO_DEVICE;
        } else {
            g->comm.assign((size_t)n, nullptr);
            const int e = nc.CommInitAll(g->comm.data(), n, g->dev.data());
            if (e != kNcclSuccess) {
                fprintf(stderr, "[treat_b200] %s\n", nccl_error("ncclCommInitAll", e).c_str());
                g->comm.clear();
                rc = TREAT_B200_ERR_CUDA;
            }
        }
    }
    g->stream.assign((size_t)n, nullptr);
    g->d_red.assign((size_t)n, nullptr);
    g->red_cap.assign((size_t)n, 0);
    for (int i = 0; i < n && rc == TREAT_B200_OK; i++) {
        if (cudaSetDevice(g->dev[i]) != cudaSuccess || cudaStreamCreateWithFlags(&g->stream[i], cudaStreamNonBlocking) != cudaSuccess) {
            cudaGetLastError();
            rc = TREAT_B200_ERR_CUDA;
        }
    }
    if (rc != TREAT_B200_OK) {
        treat_b200_group_destroy(g);
        return rc;
    }
    for (int i = 0; i < n; i++) {
        Worker *w = new Worker();
        w->start();
        g->worker.push_back(w);
    }
    *out = g;
    return TREAT_B200_OK;
}

void treat_b200_group_destroy(treat_b200_group *g)
{
    if (!g) return;
    for (Worker *w : g->worker) {
        w->shut();
        delete w;
    }
    for (size_t i = 0; i < g->comm.size(); i++)
        if (g->comm[i]) nccl().CommDestroy(g->comm[i]);
    for (int i = 0; i < g->n; i++) {
        cudaSetDevice(g->dev[i]);
        if ((size_t)i < g->d_red.size() && g->d_red[i]) cudaFree(g->d_red[i]);
        if ((size_t)i < g->stream.size() && g->stream[i]) cudaStreamDestroy(g->stream[i]);
    }
    for (treat_b200_ctx *c : g->ctx) treat_b200_destroy(c);
    delete g;
}

int32_t treat_b200_group_size(const treat_b200_group *g) { return g ? g->n : 0; }

treat_b200_ctx *treat_b200_group_ctx(treat_b200_group *g, int32_t i) { return g && i >= 0 && i < g->n ? g->ctx[i] : nullptr; }

const char *treat_b200_group_last_error(const treat_b200_group *g) { return g ? g->err.c_str() : ""; }

int treat_b200_group_set_template(treat_b200_group *g, const uint8_t *bases, int32_t m, const uint32_t *edit_site, int32_t K,
                                  const int32_t *alt_start, const int32_t *alt_end, int32_t n_alt, uint32_t edit_offset,
                                  int32_t tmpl_edit_stop, uint8_t edit_base)
{
    if (!g) return TREAT_B200_ERR_INVALID;
    int rc = on_all(g, [&](int i) {
        return treat_b200_set_template(g->ctx[i], bases, m, edit_site, K, alt_start, alt_end, n_alt, edit_offset, tmpl_edit_stop, edit_base);
    });
    g->template_set = rc == TREAT_B200_OK;
    return rc;
}

int treat_b200_group_use_template(treat_b200_group *g, const treat_b200_template *tmpl)
{
    if (!g || !tmpl) return TREAT_B200_ERR_INVALID;
    int rc = on_all(g, [&](int i) { return treat_b200_use_template(g->ctx[i], tmpl); });
    g->template_set = rc == TREAT_B200_OK;
    return rc;
}

uint32_t treat_b200_group_ops_stride(const treat_b200_group *g, uint32_t max_len) { return g ? treat_b200_ops_stride(g->ctx[0], max_len) : 0; }

int treat_b200_group_align_batch(treat_b200_group *g, const uint8_t *seqs, const uint64_t *seq_off, const uint32_t *read_count,
                                 uint64_t N, uint32_t flags, treat_b200_record *rec, uint8_t *ops, uint32_t ops_stride)
{
    if (!g || (N && (!seqs || !seq_off || !rec))) return TREAT_B200_ERR_INVALID;
    if (!g->template_set) return gfail(g, TREAT_B200_ERR_NO_TEMPLATE, "group align: no template");
    const uint64_t n = (uint64_t)g->n;
    return on_all(g, [&](int i) {
        // range i = [N i / n, N (i + 1) / n): contiguous, so the outputs land where the one-GPU call puts them
        const uint64_t lo = N * (uint64_t)i / n, hi = N * (uint64_t)(i + 1) / n;
        if (hi == lo) return (int)TREAT_B200_OK;
        return treat_b200_align_batch(g->ctx[i], seqs, seq_off + lo, read_count ? read_count + lo : nullptr, hi - lo, flags, rec + lo,
                                      ops ? ops + lo * (uint64_t)ops_stride : nullptr, ops_stride);
    });
}

int treat_b200_group_summary(treat_b200_group *g, int32_t from_rank, uint64_t *out, uint64_t len)
{
    if (!g || !out) return TREAT_B200_ERR_INVALID;
    if (!g->template_set) return gfail(g, TREAT_B200_ERR_NO_TEMPLATE, "group summary: no template");
    if (len != treat_b200_summary_len(g->ctx[0])) return gfail(g, TREAT_B200_ERR_INVALID, "group summary: len != treat_b200_summary_len");
    std::vector<uint64_t *> src((size_t)g->n, nullptr);
    for (int i = 0; i < g->n; i++) {
        int rc = treat_b200_summary_device_ptr(g->ctx[i], &src[(size_t)i]);
        if (rc) return gfail(g, rc, treat_b200_last_error(g->ctx[i]));
    }
    return group_allreduce(g, src, len, from_rank, out);
}

int treat_b200_group_heatmap(treat_b200_group *g, int32_t from_rank, uint64_t *out, uint64_t len)
{
    if (!g || !out) return TREAT_B200_ERR_INVALID;
    if (!g->template_set) return gfail(g, TREAT_B200_ERR_NO_TEMPLATE, "group heat map: no template");
    const uint64_t dim = treat_b200_heatmap_dim(g->ctx[0]);
    if (len != dim * dim) return gfail(g, TREAT_B200_ERR_INVALID, "group heat map: len != dim * dim");
    std::vector<uint64_t *> src((size_t)g->n, nullptr);
    for (int i = 0; i < g->n; i++) {
        int rc = treat_b200_heatmap_device_ptr(g->ctx[i], &src[(size_t)i]);
        if (rc) return gfail(g, rc, treat_b200_last_error(g->ctx[i]));
    }
    return group_allreduce(g, src, len, from_rank, out);
}

int treat_b200_group_reset_summary(treat_b200_group *g)
{
    if (!g) return TREAT_B200_ERR_INVALID;
    return on_all(g, [&](int i) { return treat_b200_reset_summary(g->ctx[i]); });
}

uint64_t treat_b200_group_launch_count(const treat_b200_group *g)
{
    uint64_t t = 0;
    if (g)
        for (treat_b200_ctx *c : g->ctx) t += treat_b200_launch_count(c);
    return t;
}

// ---------------------------------------------------------------------------------------------
// one process per GPU
// ---------------------------------------------------------------------------------------------
int treat_b200_comm_unique_id(uint8_t *id, size_t len)
{
    if (!id || len < TREAT_B200_COMM_ID_BYTES) return TREAT_B200_ERR_INVALID;
    Nccl &nc = nccl();
    if (!nc.ok) {
        fprintf(stderr, "[treat_b200] %s\n", nc.err.c_str());
        return TREAT_B200_ERR_NO_DEVICE;
    }
    ncclUniqueId u;
    if (nc.GetUniqueId(&u) != kNcclSuccess) return TREAT_B200_ERR_CUDA;
    memcpy(id, u.internal, TREAT_B200_COMM_ID_BYTES);
    return TREAT_B200_OK;
}

int treat_b200_comm_create(treat_b200_ctx *ctx, int32_t nranks, int32_t rank, const uint8_t *id, size_t len, treat_b200_comm **out)
{
    if (!ctx || !out || nranks < 1 || rank < 0 || rank >= nranks) return TREAT_B200_ERR_INVALID;
    *out = nullptr;
    if (nranks > 1 && (!id || len < TREAT_B200_COMM_ID_BYTES)) return TREAT_B200_ERR_INVALID;
    treat_b200_comm *c = new treat_b200_comm();
    c->ctx = ctx;
    c->device = ctx_device(ctx);
    c->nranks = nranks;
    c->rank = rank;
    if (cudaSetDevice(c->device) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError();
        delete c;
        return TREAT_B200_ERR_CUDA;
    }
    if (nranks > 1) {
        Nccl &nc = nccl();
        if (!nc.ok) {
            fprintf(stderr, "[treat_b200] %s\n", nc.err.c_str());
            treat_b200_comm_destroy(c);
            return TREAT_B200_ERR_NO_DEVICE;
        }
        ncclUniqueId u;
        memcpy(u.internal, id, TREAT_B200_COMM_ID_BYTES);
        const int e = nc.CommInitRank(&c->comm, nranks, u, rank);
        if (e != kNcclSuccess) {
            fprintf(stderr, "[treat_b200] %s\n", nccl_error("ncclCommInitRank", e).c_str());
            c->comm = nullptr;
            treat_b200_comm_destroy(c);
            return TREAT_B200_ERR_CUDA;
        }
    }
    *out = c;
    return TREAT_B200_OK;
}

void treat_b200_comm_destroy(treat_b200_comm *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm) nccl().CommDestroy(c->comm);
    if (c->d_red) cudaFree(c->d_red);
    for (int i = 0; i < 2; i++) {
        if (c->d_snap[i]) cudaFree(c->d_snap[i]);
        if (c->snap_done[i]) cudaEventDestroy(c->snap_done[i]);
        if (c->red_done[i]) cudaEventDestroy(c->red_done[i]);
    }
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

const char *treat_b200_comm_last_error(const treat_b200_comm *c) { return c ? c->err.c_str() : ""; }

#define CCK(c, expr)                                                          \
    do {                                                                      \
        cudaError_t _e = (expr);                                              \
        if (_e != cudaSuccess) {                                              \
            (c)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);    \
            return TREAT_B200_ERR_CUDA;                                       \
        }                                                                     \
    } while (0)

// The reduction runs on the communicator's own stream over a snapshot of the counters taken on the caller's stream, so the
// caller's next align call does not wait for the other ranks: a rank that is late delays the reduction, not the kernels.
int treat_b200_comm_reduce_summary_device(treat_b200_comm *c, uint64_t *d_out, void *stream)
{
    if (!c || !d_out) return TREAT_B200_ERR_INVALID;
    uint64_t *src = nullptr;
    int rc = treat_b200_summary_device_ptr(c->ctx, &src);
    if (rc) return rc;
    const uint64_t len = treat_b200_summary_len(c->ctx);
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx_stream(c->ctx);
    CCK(c, cudaSetDevice(c->device));
    const int i = (int)(c->turn++ & 1u);
    if (!c->snap_done[i]) {
        CCK(c, cudaEventCreateWithFlags(&c->snap_done[i], cudaEventDisableTiming));
        CCK(c, cudaEventCreateWithFlags(&c->red_done[i], cudaEventDisableTiming));
    }
    if (c->red_pending[i]) CCK(c, cudaStreamWaitEvent(st, c->red_done[i], 0));  // the reduction that last read this snapshot
    if (c->snap_cap[i] < len * 8) {
        CCK(c, cudaStreamSynchronize(c->stream));
        if (c->d_snap[i]) cudaFree(c->d_snap[i]);
        c->d_snap[i] = nullptr;
        c->snap_cap[i] = 0;
        CCK(c, cudaMalloc((void **)&c->d_snap[i], len * 8));
        c->snap_cap[i] = len * 8;
    }
    CCK(c, cudaMemcpyAsync(c->d_snap[i], src, len * 8, cudaMemcpyDeviceToDevice, st));
    CCK(c, cudaEventRecord(c->snap_done[i], st));
    CCK(c, cudaStreamWaitEvent(c->stream, c->snap_done[i], 0));
    if (c->nranks == 1) {
        CCK(c, cudaMemcpyAsync(d_out, c->d_snap[i], len * 8, cudaMemcpyDeviceToDevice, c->stream));
    } else {
        const int e = nccl().AllReduce(c->d_snap[i], d_out, (size_t)len, kNcclUint64, kNcclSum, c->comm, c->stream);
        if (e != kNcclSuccess) {
            c->err = nccl_error("ncclAllReduce", e);
            return TREAT_B200_ERR_CUDA;
        }
    }
    CCK(c, cudaEventRecord(c->red_done[i], c->stream));
    c->red_pending[i] = true;
    return TREAT_B200_OK;
}

int treat_b200_comm_join(treat_b200_comm *c, void *stream)
{
    if (!c) return TREAT_B200_ERR_INVALID;
    CCK(c, cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx_stream(c->ctx);
    const int last = (int)((c->turn + 1u) & 1u);  // the most recent reduction; the one before it precedes it on the same stream
    if (c->turn && c->red_pending[last]) CCK(c, cudaStreamWaitEvent(st, c->red_done[last], 0));
    return TREAT_B200_OK;
}

static int comm_reduce_host(treat_b200_comm *c, uint64_t *src, uint64_t len, uint64_t *out)
{
    int rc = treat_b200_synchronize(c->ctx);  // the counters are final
    if (rc) return rc;
    if (cudaSetDevice(c->device) != cudaSuccess) return TREAT_B200_ERR_CUDA;
    if ((rc = reserve_red(c->d_red, c->red_cap, len * 8))) return rc;
    if (c->nranks == 1) {
        if (cudaMemcpy(out, src, len * 8, cudaMemcpyDeviceToHost) != cudaSuccess) {
            c->err = cudaGetErrorString(cudaGetLastError());
            return TREAT_B200_ERR_CUDA;
        }
        return TREAT_B200_OK;
    }
    const int e = nccl().AllReduce(src, c->d_red, (size_t)len, kNcclUint64, kNcclSum, c->comm, c->stream);
    if (e != kNcclSuccess) {
        c->err = nccl_error("ncclAllReduce", e);
        return TREAT_B200_ERR_CUDA;
    }
    if (cudaMemcpyAsync(out, c->d_red, len * 8, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
        cudaStreamSynchronize(c->stream) != cudaSuccess) {
        c->err = cudaGetErrorString(cudaGetLastError());
        return TREAT_B200_ERR_CUDA;
    }
    return TREAT_B200_OK;
}

int treat_b200_comm_reduce_summary(treat_b200_comm *c, uint64_t *out, uint64_t len)
{
    if (!c || !out) return TREAT_B200_ERR_INVALID;
    if (len != treat_b200_summary_len(c->ctx)) return TREAT_B200_ERR_INVALID;
    uint64_t *src = nullptr;
    int rc = treat_b200_summary_device_ptr(c->ctx, &src);
    if (rc) return rc;
    return comm_reduce_host(c, src, len, out);
}

int treat_b200_comm_reduce_heatmap(treat_b200_comm *c, uint64_t *out, uint64_t len)
{
    if (!c || !out) return TREAT_B200_ERR_INVALID;
    const uint64_t dim = treat_b200_heatmap_dim(c->ctx);
    if (len != dim * dim) return TREAT_B200_ERR_INVALID;
    uint64_t *src = nullptr;
    int rc = treat_b200_heatmap_device_ptr(c->ctx, &src);
    if (rc) return rc;
    return comm_reduce_host(c, src, len, out);
}

}  // extern "C"
