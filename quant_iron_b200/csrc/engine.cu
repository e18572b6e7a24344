// engine.cu -- context, error reporting, options, kernel accounting.
#include <stdarg.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace qi {

void jit_drain();      // window.cu (tile_jit.cuh)
void jit_stats(uint64_t*, uint64_t*, uint64_t*, double*, double*, int);

const char* const kFamilyNames[KF_COUNT] = {
    "init", "gate_pair", "gate_diag", "gate_swap", "gate_matchgate", "gate_window", "pauli_apply",
    "pauli_exp", "pauli_expect", "reduce", "elementwise", "probabilities", "scan", "sample",
    "collapse", "exchange", "barrier", "pauli_exp_window", "gate_tile", "gate_tile_jit"};

static thread_local uint64_t tl_payload[2] = {0, 0};
static thread_local char tl_msg[256] = {0};

void set_error(uint64_t p0, uint64_t p1, const char* fmt, ...) {
    tl_payload[0] = p0;
    tl_payload[1] = p1;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(tl_msg, sizeof(tl_msg), fmt, ap);
    va_end(ap);
}

int fail(int status, uint64_t p0, uint64_t p1, const char* msg) {
    set_error(p0, p1, "%s", msg);
    return status;
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error((uint64_t)e, 0, "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    cudaGetLastError();  // clear sticky-less errors
    return QI_ERR_CUDA;
}

static Context g_ctx;
static std::mutex g_mutex;

Context& ctx() { return g_ctx; }

static int init_locked(int device) {
    Context& c = g_ctx;
    if (c.ready && (device < 0 || device == c.device)) return QI_OK;
    if (c.ready)       // the stream, events, scratch and pooled buffers live on the first device: re-binding would mix devices
        return fail(QI_ERR_INVALID_ARGUMENT, (uint64_t)device, (uint64_t)c.device, "the engine is already bound to another device (one process per GPU)");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error((uint64_t)e, 0, "no CUDA device available (%s): libqiron_b200 has no CPU fallback",
                  cudaGetErrorString(e));
        cudaGetLastError();
        return QI_ERR_CUDA;
    }
    if (device < 0) {
        QI_CUDA(cudaGetDevice(&device));
    }
    QI_CUDA(cudaSetDevice(device));
    c.device = device;
    cudaDeviceProp prop;
    QI_CUDA(cudaGetDeviceProperties(&prop, device));
    c.sm_count = prop.multiProcessorCount;
    if (!c.stream) QI_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    if (!c.d_result) QI_CUDA(cudaMalloc(&c.d_result, 4096 * sizeof(double)));
    if (!c.h_result) QI_CUDA(cudaMallocHost(&c.h_result, 4096 * sizeof(double)));
    if (!c.timer_start) QI_CUDA(cudaEventCreate(&c.timer_start));
    if (!c.timer_stop) QI_CUDA(cudaEventCreate(&c.timer_stop));
    c.ready = true;
    return QI_OK;
}

int ensure_ctx() {
    if (g_ctx.ready) return cudaSetDevice(g_ctx.device) == cudaSuccess ? QI_OK : cuda_fail(cudaGetLastError(), "cudaSetDevice");
    std::lock_guard<std::mutex> lk(g_mutex);
    return init_locked(-1);
}

// ---- device-buffer cache ---------------------------------------------------------------------------
namespace {
struct DevPool {
    std::mutex mu;
    std::unordered_map<size_t, std::vector<void*>> free_by_size;
    size_t cached_bytes = 0;
};
DevPool& pool() { static DevPool p; return p; }
const size_t kMaxPooledBuffer = 1ull << 30;      // larger buffers go straight back to the driver
}  // namespace

int dev_alloc(void** p, size_t bytes) {
    *p = nullptr;
    if (bytes == 0) return QI_OK;
    {
        DevPool& dp = pool();
        std::lock_guard<std::mutex> lk(dp.mu);
        auto it = dp.free_by_size.find(bytes);
        if (it != dp.free_by_size.end() && !it->second.empty()) {
            *p = it->second.back();
            it->second.pop_back();
            dp.cached_bytes -= bytes;
            return QI_OK;
        }
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) {
        dev_pool_trim();                             // give the cache back and retry once
        cudaGetLastError();
        e = cudaMalloc(p, bytes);
    }
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    return QI_OK;
}

void dev_free(void* p, size_t bytes) {
    if (!p) return;
    Context& c = ctx();
    const size_t cap = (size_t)(c.opt_pool_mb > 0 ? c.opt_pool_mb : 0) << 20;
    if (bytes <= kMaxPooledBuffer) {
        DevPool& dp = pool();
        std::lock_guard<std::mutex> lk(dp.mu);
        if (dp.cached_bytes + bytes <= cap) {
            // work queued on the engine stream may still touch the buffer; the next user is queued behind it
            dp.free_by_size[bytes].push_back(p);
            dp.cached_bytes += bytes;
            return;
        }
    }
    if (c.ready) cudaStreamSynchronize(c.stream);
    cudaFree(p);
}

void dev_pool_trim() {
    DevPool& dp = pool();
    std::lock_guard<std::mutex> lk(dp.mu);
    if (ctx().ready) cudaStreamSynchronize(ctx().stream);
    for (auto& kv : dp.free_by_size)
        for (void* p : kv.second) cudaFree(p);
    dp.free_by_size.clear();
    dp.cached_bytes = 0;
}

int ensure_partials(size_t doubles) {
    Context& c = g_ctx;
    if (c.partial_capacity >= doubles) return QI_OK;
    if (c.d_partials) {
        QI_CUDA(cudaStreamSynchronize(c.stream));
        QI_CUDA(cudaFree(c.d_partials));
        c.d_partials = nullptr;
    }
    size_t cap = doubles < (1u << 16) ? (1u << 16) : doubles;
    QI_CUDA(cudaMalloc(&c.d_partials, cap * sizeof(double)));
    c.partial_capacity = cap;
    return QI_OK;
}

int grid_for(uint64_t work_items, int block, int max_waves) {
    uint64_t blocks = (work_items + block - 1) / block;
    uint64_t cap = (uint64_t)g_ctx.sm_count * (2048 / block) * max_waves;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

LaunchScope::LaunchScope(int fam, double bytes) : family(fam) {
    Context& c = g_ctx;
    c.launches[fam]++;
    c.alg_bytes[fam] += bytes;
    if (c.opt_profile) {
        auto get = [&]() {
            cudaEvent_t ev;
            if (!c.event_pool.empty()) { ev = c.event_pool.back(); c.event_pool.pop_back(); }
            else cudaEventCreate(&ev);
            return ev;
        };
        start = get();
        stop = get();
        cudaEventRecord(start, c.stream);
    }
}

LaunchScope::~LaunchScope() {
    Context& c = g_ctx;
    if (start) {
        cudaEventRecord(stop, c.stream);
        c.pending.push_back({start, stop, family});
    }
}

static void drain_profile() {
    Context& c = g_ctx;
    if (c.pending.empty()) return;
    cudaStreamSynchronize(c.stream);
    for (auto& p : c.pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.start, p.stop) == cudaSuccess) c.total_ms[p.family] += ms;
        c.event_pool.push_back(p.start);
        c.event_pool.push_back(p.stop);
    }
    c.pending.clear();
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, what);
    return QI_OK;
}

}  // namespace qi

using namespace qi;

extern "C" {

void qi_last_error(uint64_t payload[2], char* msg, size_t msg_len) {
    if (payload) { payload[0] = tl_payload[0]; payload[1] = tl_payload[1]; }
    if (msg && msg_len) { strncpy(msg, tl_msg, msg_len - 1); msg[msg_len - 1] = 0; }
}

const char* qi_version(void) { return "qiron_b200 0.1 (sm_100a)"; }

int qi_init(int device) {
    std::lock_guard<std::mutex> lk(g_mutex);
    return init_locked(device);
}

int qi_jit_drain(void) { qi::jit_drain(); return QI_OK; }
int qi_jit_stats(uint64_t* modules, uint64_t* failed, uint64_t* pending, double* assemble_ms, double* fp64_warp_instr) {
    uint64_t a = 0, b = 0, c = 0;
    double d = 0.0, e = 0.0;
    qi::jit_stats(&a, &b, &c, &d, &e, 0);
    if (fp64_warp_instr) *fp64_warp_instr = e;
    if (modules) *modules = a;
    if (failed) *failed = b;
    if (pending) *pending = c;
    if (assemble_ms) *assemble_ms = d;
    return QI_OK;
}

int qi_synchronize(void) {
    QI_TRY(ensure_ctx());
    QI_CUDA(cudaStreamSynchronize(ctx().stream));
    return QI_OK;
}

int qi_device_info(char* name, size_t name_len, int* sm_count, uint64_t* total_mem, uint64_t* free_mem) {
    QI_TRY(ensure_ctx());
    cudaDeviceProp prop;
    QI_CUDA(cudaGetDeviceProperties(&prop, ctx().device));
    if (name && name_len) { strncpy(name, prop.name, name_len - 1); name[name_len - 1] = 0; }
    if (sm_count) *sm_count = prop.multiProcessorCount;
    size_t f = 0, t = 0;
    QI_CUDA(cudaMemGetInfo(&f, &t));
    if (total_mem) *total_mem = t;
    if (free_mem) *free_mem = f;
    return QI_OK;
}

int qi_set_option(const char* name, int64_t value) {
    if (!name) return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "option name is NULL");
    Context& c = ctx();
    if (!strcmp(name, "path")) c.opt_path = (int)value;
    else if (!strcmp(name, "fuse")) c.opt_fuse = (int)value;
    else if (!strcmp(name, "window_regs")) c.opt_window_regs = (int)value;
    else if (!strcmp(name, "lazy_swap")) c.opt_lazy_swap = (int)value;
    else if (!strcmp(name, "tma")) c.opt_tma = (int)value;
    else if (!strcmp(name, "absorb")) c.opt_absorb = (int)value;
    else if (!strcmp(name, "lean")) c.opt_lean = (int)value;
    else if (!strcmp(name, "tile")) c.opt_tile = (int)value;
    else if (!strcmp(name, "jit")) c.opt_jit = (int)value;
    else if (!strcmp(name, "jit_min_qubits")) c.opt_jit_min_qubits = (int)value;
    else if (!strcmp(name, "jit_ctas")) c.opt_jit_ctas = (int)value;
    else if (!strcmp(name, "jit_groups")) c.opt_jit_groups = (int)value;
    else if (!strcmp(name, "tile_lean")) c.opt_tile_lean = (int)value;
    else if (!strcmp(name, "tile_pform")) c.opt_tile_pform = (int)value;
    else if (!strcmp(name, "tile_carry")) c.opt_tile_carry = (int)value;
    else if (!strcmp(name, "tile_restore")) c.opt_tile_restore = (int)value;
    else if (!strcmp(name, "pauli_unit")) c.opt_pauli_unit = (int)value;
    else if (!strcmp(name, "tile_bfs_by_use")) c.opt_tile_bfs_by_use = (int)value;
    else if (!strcmp(name, "jit_prefetch")) c.opt_jit_prefetch = (int)value;
    else if (!strcmp(name, "jit_stage")) c.opt_jit_stage = (int)value;
    else if (!strcmp(name, "jit_smem_kb")) c.opt_jit_smem_kb = (int)value;
    else if (!strcmp(name, "debug_ptx")) c.opt_debug_ptx = (int)value;
    else if (!strcmp(name, "tile_slide")) c.opt_tile_slide = (int)value;
    else if (!strcmp(name, "tile_absorb")) c.opt_tile_absorb = (int)value;
    else if (!strcmp(name, "peer_timeout_s")) c.opt_peer_timeout_s = (int)value;
    else if (!strcmp(name, "cz_rewrite")) c.opt_cz_rewrite = (int)value;
    else if (!strcmp(name, "tile_min_qubits")) c.opt_tile_min_qubits = (int)value;
    else if (!strcmp(name, "tile_min_gates")) c.opt_tile_min_gates = (int)value;
    else if (!strcmp(name, "prefetch")) c.opt_prefetch = (int)value;
    else if (!strcmp(name, "late_tables")) c.opt_late_tables = (int)value;
    else if (!strcmp(name, "host_chunk_qubits")) c.opt_host_chunk_qubits = (int)(value < 0 ? 0 : (value > 4 ? 4 : value));
    else if (!strcmp(name, "host_min_qubits")) c.opt_host_min_qubits = (int)value;
    else if (!strcmp(name, "pool_mb")) { c.opt_pool_mb = value; if (value <= 0) dev_pool_trim(); }
    else if (!strcmp(name, "profile")) { if (!value) drain_profile(); c.opt_profile = (int)value; }
    else return fail(QI_ERR_INVALID_ARGUMENT, 0, 0, "unknown option");
    return QI_OK;
}

int qi_stats_reset(void) {
    Context& c = ctx();
    drain_profile();
    for (int i = 0; i < KF_COUNT; i++) { c.launches[i] = 0; c.alg_bytes[i] = 0; c.total_ms[i] = 0; }
    uint64_t a, b, d;
    double e, f;
    qi::jit_stats(&a, &b, &d, &e, &f, 1);          // the FP64 instruction count of the JIT modules restarts with the launch counters
    return QI_OK;
}

int qi_stats_get(qi_kernel_stat* out, int capacity, int* count) {
    Context& c = ctx();
    drain_profile();
    int n = 0;
    for (int i = 0; i < KF_COUNT && n < capacity; i++) {
        if (!c.launches[i]) continue;
        memset(&out[n], 0, sizeof(out[n]));
        strncpy(out[n].name, kFamilyNames[i], sizeof(out[n].name) - 1);
        out[n].launches = c.launches[i];
        out[n].total_ms = c.total_ms[i];
        out[n].algorithmic_bytes = c.alg_bytes[i];
        n++;
    }
    if (count) *count = n;
    return QI_OK;
}

int qi_timer_start(void) {
    QI_TRY(ensure_ctx());
    QI_CUDA(cudaEventRecord(ctx().timer_start, ctx().stream));
    return QI_OK;
}

int qi_timer_stop(float* elapsed_ms) {
    QI_TRY(ensure_ctx());
    QI_CUDA(cudaEventRecord(ctx().timer_stop, ctx().stream));
    QI_CUDA(cudaEventSynchronize(ctx().timer_stop));
    float ms = 0.f;
    QI_CUDA(cudaEventElapsedTime(&ms, ctx().timer_start, ctx().timer_stop));
    if (elapsed_ms) *elapsed_ms = ms;
    return QI_OK;
}

double qi_uniform(uint64_t seed, uint64_t k) { return uniform_at(seed, k); }

}  // extern "C"
