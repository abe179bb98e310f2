// C ABI of libvpb200 (include/vpb200.h): context, pooled buffers, host-buffer stage calls, pipeline.
//
// Replaces the reference's per-call CudaPtr<T> malloc/copy/free pattern (vplib/src/cuda_ptr.h:14-93,
// e.g. vox/naive.cu:91-121, jfa/tiled.cu:250-336) with one stream, grow-only device buffers that survive
// between calls, and stage-to-stage device residency inside vpb_pipeline_host.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace vpb {

namespace {

thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
std::mutex g_api_mutex;          // vpb_init / vpb_shutdown; the compute entry points keep vplib's one-caller-thread rule

struct Buf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return VPB_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        // round up so that repeated slightly-growing requests do not reallocate
        size_t want = ((bytes + (1u << 20) - 1) >> 20) << 20;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("device allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
            p = nullptr;
            return VPB_ERR_NOMEM;
        }
        cap = want;
        return VPB_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T* as() const { return static_cast<T*>(p); }
};

struct Context {
    bool ready = false;
    int device = 0;
    int sms = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;           // D2H of finished z-chunks of the sdf while the final pass still runs
    cudaEvent_t chunk_ev[16] = {};
    cudaEvent_t copy_done = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    float timing[3] = {0, 0, 0};
    Buf verts, tris, grid_a, grid_b, grid_c, scratch, state_a, state_b, sdf, seeds;
    // asynchronous pipeline (vpb_pipeline_submit / vpb_pipeline_wait): two jobs in flight, each slot owns the device
    // copies of its results so that the next job's kernels can run while this one's D2H is still going
    Buf slot_words[2], slot_sdf[2];
    cudaEvent_t slot_done[2] = {nullptr, nullptr};
    uint64_t slot_ticket[2] = {0, 0};             // ticket whose results the slot holds (0 = none)
    uint64_t next_ticket = 1;
};
Context g_ctx;

#define VPB_TRY(expr)                 \
    do {                              \
        int _rc = (expr);             \
        if (_rc != VPB_OK) return _rc; \
    } while (0)

int require_ready() {
    if (!g_ctx.ready) {
        set_error("vpb_init has not been called (or failed)");
        return VPB_ERR_STATE;
    }
    return VPB_OK;
}

Frame make_frame(uint32_t n, float vs, const float origin[3]) { return Frame{origin[0], origin[1], origin[2], vs, n}; }

// ---- state width dispatch: 32-bit words up to N = 1024, 64-bit up to 2048 (common.cuh) ----------------------------
constexpr uint32_t kMaxJfaN = 2048;
size_t state_size(uint32_t n) { return jfa_state64(n) ? 8 : 4; }
int seed_any(const uint32_t* words, uint32_t n, uint32_t z0, uint32_t z1, uint32_t* state, cudaStream_t st) {
    return jfa_state64(n) ? jfa_seed_launch_s64(words, n, z0, z1, state, st) : jfa_seed_launch(words, n, z0, z1, state, st);
}
int early_any(const uint32_t* words, const Frame& f, uint32_t z0, uint32_t z1, uint32_t* shell_scratch, uint32_t* state,
              cudaStream_t st, bool shell_ready = false) {
    if (shell_ready)
        return jfa_state64(f.n) ? jfa_early_from_shell_launch_s64(shell_scratch, f, z0, z1, state, st)
                                : jfa_early_from_shell_launch(shell_scratch, f, z0, z1, state, st);
    return jfa_state64(f.n) ? jfa_early_launch_s64(words, f, z0, z1, shell_scratch, state, st)
                            : jfa_early_launch(words, f, z0, z1, shell_scratch, state, st);
}
bool early_takes(const Frame& f) { return (jfa_state64(f.n) ? jfa_early_supported_s64(f) : jfa_early_supported(f)) != 0; }
int pass_any(const uint32_t* below, const uint32_t* mid, const uint32_t* above, uint32_t* dst, const Frame& f, uint32_t z0,
             uint32_t z1, uint32_t k, const uint32_t* words, float* sdf, uint32_t* seeds, cudaStream_t st) {
    return jfa_state64(f.n) ? jfa_pass_launch_s64(below, mid, above, dst, f, z0, z1, k, words, sdf, seeds, st)
                            : jfa_pass_launch(below, mid, above, dst, f, z0, z1, k, words, sdf, seeds, st);
}
int finalize_any(const uint32_t* state, const Frame& f, uint32_t z0, uint32_t z1, const uint32_t* words, float* sdf,
                 uint32_t* seeds, cudaStream_t st) {
    return jfa_state64(f.n) ? jfa_finalize_launch_s64(state, f, z0, z1, words, sdf, seeds, st)
                            : jfa_finalize_launch(state, f, z0, z1, words, sdf, seeds, st);
}

// Where the host-buffer calls want the result: the final pass is then issued in z-chunks and every finished chunk of the
// signed distance (and of the seed indices) starts its D2H copy on a second stream while the next chunk is computed.
struct HostSink {
    float* sdf_host = nullptr;
    uint32_t* seeds_host = nullptr;
    cudaEvent_t kernels_done = nullptr;           // recorded on the compute stream after the last chunk's kernel
    cudaEvent_t copies_done = nullptr;            // non-null: recorded on the copy stream after the last copy and the
                                                  // compute stream does NOT wait for the copies (asynchronous jobs)
};

// Runs seed extraction + all passes + signed output on one GPU.  sdf may be NULL: the final pass never writes its
// state destination, so the signed distance then goes INTO that free state buffer (4 B/voxel of HBM saved; at 2048^3
// that is what makes the job fit one GPU) and *sdf_at tells the caller where it is.
// shell_ready: the seed-shell bits of `words` are already in sb (the fused CSG + shell kernel wrote them).
int jfa_run(const uint32_t* words, const Frame& f, uint32_t* sa, uint32_t* sb, float* sdf, uint32_t* seeds, cudaStream_t st,
            float** sdf_at = nullptr, const HostSink* sink = nullptr, bool shell_ready = false) {
    const uint32_t n = f.n;
    // seed extraction + the passes k = N/2, N/4, N/8 fused (jfa_early.cu; the shell bits go through the free buffer sb),
    // or, for shapes it does not take, seed extraction and every pass on its own
    uint32_t k_first = n / 2;
    {
        const int rc = early_any(words, f, 0, n, sb, sa, st, shell_ready);
        if (rc < 0) return rc;
        if (rc == 0) k_first = n / 16;
        else VPB_TRY(seed_any(words, n, 0, n, sa, st));
    }
    const uint64_t plane_bytes = (uint64_t)n * n * state_size(n);
    char* in = reinterpret_cast<char*>(sa);
    char* out = reinterpret_cast<char*>(sb);
    if (n / 2 == 0) {
        float* target = sdf ? sdf : reinterpret_cast<float*>(out);
        if (sdf_at) *sdf_at = target;
        return finalize_any(sa, f, 0, n, words, target, seeds, st);
    }
    for (uint32_t k = k_first; k >= 1; k /= 2) {
        const bool last = (k == 1);
        float* target = sdf ? sdf : reinterpret_cast<float*>(out);
        if (last && sdf_at) *sdf_at = target;
        if (last && sink && n >= 256) {
            // final pass in z-chunks; chunk c's D2H overlaps chunk c+1's kernel (the copies are all enqueued after the
            // kernels, so a pageable destination, whose copies block the host, cannot delay a launch)
            const uint32_t chunks = 8, tc = (n + chunks - 1) / chunks;
            const size_t plane_vox = (size_t)n * n;
            uint32_t used = 0;
            for (uint32_t z0 = 0; z0 < n; z0 += tc, ++used) {
                const uint32_t z1 = z0 + tc < n ? z0 + tc : n;
                char* mid = in + (uint64_t)z0 * plane_bytes;
                VPB_TRY(pass_any(reinterpret_cast<uint32_t*>(mid - plane_bytes), reinterpret_cast<uint32_t*>(mid),
                                 reinterpret_cast<uint32_t*>(mid + plane_bytes),
                                 reinterpret_cast<uint32_t*>(out + (uint64_t)z0 * plane_bytes), f, z0, z1, 1, words,
                                 target + z0 * plane_vox, seeds ? seeds + z0 * plane_vox : nullptr, st));
                VPB_CUDA(cudaEventRecord(g_ctx.chunk_ev[used], st));
            }
            if (sink->kernels_done) VPB_CUDA(cudaEventRecord(sink->kernels_done, st));
            used = 0;
            for (uint32_t z0 = 0; z0 < n; z0 += tc, ++used) {
                const uint32_t z1 = z0 + tc < n ? z0 + tc : n;
                const size_t off = z0 * plane_vox, cnt = (size_t)(z1 - z0) * plane_vox;
                VPB_CUDA(cudaStreamWaitEvent(g_ctx.copy_stream, g_ctx.chunk_ev[used], 0));
                VPB_CUDA(cudaMemcpyAsync(sink->sdf_host + off, target + off, cnt * 4, cudaMemcpyDeviceToHost, g_ctx.copy_stream));
                if (seeds && sink->seeds_host)
                    VPB_CUDA(cudaMemcpyAsync(sink->seeds_host + off, seeds + off, cnt * 4, cudaMemcpyDeviceToHost, g_ctx.copy_stream));
            }
            if (sink->copies_done) {
                VPB_CUDA(cudaEventRecord(sink->copies_done, g_ctx.copy_stream));
            } else {
                VPB_CUDA(cudaEventRecord(g_ctx.copy_done, g_ctx.copy_stream));
                VPB_CUDA(cudaStreamWaitEvent(st, g_ctx.copy_done, 0));
            }
            return VPB_OK;
        }
        VPB_TRY(pass_any(reinterpret_cast<uint32_t*>(in - k * plane_bytes), reinterpret_cast<uint32_t*>(in),
                         reinterpret_cast<uint32_t*>(in + k * plane_bytes), reinterpret_cast<uint32_t*>(out), f, 0, n, k, words,
                         last ? target : nullptr, last ? seeds : nullptr, st));
        char* t = in; in = out; out = t;
    }
    return VPB_OK;
}

}  // namespace

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
void count_launch(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int num_sms() { return g_ctx.sms; }

}  // namespace vpb

using namespace vpb;

extern "C" {

int vpb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

static int create_resources() {
    VPB_CUDA(cudaStreamCreateWithFlags(&g_ctx.stream, cudaStreamNonBlocking));
    VPB_CUDA(cudaStreamCreateWithFlags(&g_ctx.copy_stream, cudaStreamNonBlocking));
    for (auto& ev : g_ctx.ev) VPB_CUDA(cudaEventCreate(&ev));
    for (auto& ev : g_ctx.chunk_ev) VPB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    VPB_CUDA(cudaEventCreateWithFlags(&g_ctx.copy_done, cudaEventDisableTiming));
    for (auto& ev : g_ctx.slot_done) VPB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    return VPB_OK;
}

// frees whatever exists (also the partially built context of a failed vpb_init)
static void release_resources() {
    if (g_ctx.stream) cudaStreamSynchronize(g_ctx.stream);
    if (g_ctx.copy_stream) cudaStreamSynchronize(g_ctx.copy_stream);
    for (Buf* b : {&g_ctx.verts, &g_ctx.tris, &g_ctx.grid_a, &g_ctx.grid_b, &g_ctx.grid_c, &g_ctx.scratch, &g_ctx.state_a,
                   &g_ctx.state_b, &g_ctx.sdf, &g_ctx.seeds, &g_ctx.slot_words[0], &g_ctx.slot_words[1], &g_ctx.slot_sdf[0],
                   &g_ctx.slot_sdf[1]})
        b->release();
    jfa_lut_release();
    jfa_lut_release_s64();
    for (auto& ev : g_ctx.ev) { if (ev) cudaEventDestroy(ev); ev = nullptr; }
    for (auto& ev : g_ctx.chunk_ev) { if (ev) cudaEventDestroy(ev); ev = nullptr; }
    for (auto& ev : g_ctx.slot_done) { if (ev) cudaEventDestroy(ev); ev = nullptr; }
    if (g_ctx.copy_done) cudaEventDestroy(g_ctx.copy_done);
    g_ctx.copy_done = nullptr;
    if (g_ctx.copy_stream) cudaStreamDestroy(g_ctx.copy_stream);
    g_ctx.copy_stream = nullptr;
    if (g_ctx.stream) cudaStreamDestroy(g_ctx.stream);
    g_ctx.stream = nullptr;
}

int vpb_init(int device) {
    std::lock_guard<std::mutex> guard(g_api_mutex);
    if (g_ctx.ready && g_ctx.device == device) return VPB_OK;
    if (g_ctx.ready) {
        cudaSetDevice(g_ctx.device);
        release_resources();
        g_ctx.ready = false;
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        set_error("no CUDA device available (%s); vpb200 has no CPU fallback", e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
        return VPB_ERR_CUDA;
    }
    VPB_REQUIRE(device >= 0 && device < count, "vpb_init: device %d out of range (count %d)", device, count);
    VPB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    VPB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        return VPB_ERR_CUDA;
    }
    g_ctx.sms = prop.multiProcessorCount;
    g_ctx.device = device;
    const int rc = create_resources();
    if (rc != VPB_OK) {
        release_resources();
        return rc;
    }
    g_ctx.slot_ticket[0] = g_ctx.slot_ticket[1] = 0;
    g_ctx.next_ticket = 1;
    g_ctx.ready = true;
    g_launches = 0;
    return VPB_OK;
}

void vpb_shutdown(void) {
    std::lock_guard<std::mutex> guard(g_api_mutex);
    if (!g_ctx.ready) return;
    cudaSetDevice(g_ctx.device);
    release_resources();
    g_ctx.ready = false;
}

const char* vpb_last_error(void) { return g_err; }
uint64_t vpb_kernel_launches(void) { return g_launches.load(std::memory_order_relaxed); }

int vpb_last_timing(float out[3]) {
    VPB_REQUIRE(out, "vpb_last_timing: null");
    memcpy(out, g_ctx.timing, sizeof g_ctx.timing);
    return VPB_OK;
}

int vpb_fnv1a64_chunks(const void* data, uint64_t bytes, uint32_t chunks, uint64_t* out) {
    VPB_REQUIRE(out && chunks >= 1 && chunks <= 4096 && (data || bytes == 0), "fnv1a64_chunks: bad argument");
    VPB_REQUIRE(bytes % chunks == 0, "fnv1a64_chunks: %llu bytes do not split into %u equal chunks", (unsigned long long)bytes, chunks);
    const uint64_t per = bytes / chunks;
    auto one = [=](uint32_t c) {
        const unsigned char* p = static_cast<const unsigned char*>(data) + (uint64_t)c * per;
        uint64_t h = 1469598103934665603ull;
        for (uint64_t i = 0; i < per; ++i) { h ^= p[i]; h *= 1099511628211ull; }
        out[c] = h;
    };
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const uint32_t nthreads = std::min<uint32_t>(chunks, std::min<unsigned>(hw, 64u));
    if (nthreads <= 1) {
        for (uint32_t c = 0; c < chunks; ++c) one(c);
        return VPB_OK;
    }
    std::vector<std::thread> pool;
    for (uint32_t t = 0; t < nthreads; ++t)
        pool.emplace_back([=] { for (uint32_t c = t; c < chunks; c += nthreads) one(c); });
    for (auto& th : pool) th.join();
    return VPB_OK;
}

// ---------------------------------------------------------------------------------- device-pointer calls

size_t vpb_voxelize_scratch_bytes(uint32_t n, uint64_t n_tris, uint32_t z0, uint32_t z1) {
    return vox_scratch_bytes(n, n_tris, z0, z1);
}

int vpb_voxelize_dev(const float* verts, uint64_t n_verts, const uint32_t* tris, uint64_t n_tris, uint32_t n, float vs,
                     const float origin[3], uint32_t z0, uint32_t z1, uint32_t* words_slab, void* scratch,
                     size_t scratch_bytes, void* stream) {
    VPB_TRY(require_ready());
    VPB_REQUIRE(origin, "voxelize: null origin");
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : g_ctx.stream;
    return vox_launch(verts, n_verts, tris, n_tris, make_frame(n, vs, origin), z0, z1, words_slab, scratch, scratch_bytes, st);
}

size_t vpb_voxelize_surface_scratch_bytes(uint64_t n_tris) { return vox_surface_scratch_bytes(n_tris); }

int vpb_voxelize_surface_dev(const float* verts, uint64_t n_verts, const uint32_t* tris, uint64_t n_tris, uint32_t n, float vs,
                             const float origin[3], uint32_t z0, uint32_t z1, uint32_t* words_slab, void* scratch,
                             size_t scratch_bytes, void* stream) {
    VPB_TRY(require_ready());
    VPB_REQUIRE(origin, "voxelize_surface: null origin");
    return vox_surface_launch(verts, n_verts, tris, n_tris, make_frame(n, vs, origin), z0, z1, words_slab, scratch, scratch_bytes,
                              stream ? static_cast<cudaStream_t>(stream) : g_ctx.stream);
}

int vpb_csg_dev(uint32_t* a, const uint32_t* b, uint64_t n_words, int op, void* stream) {
    VPB_TRY(require_ready());
    return csg_launch(a, b, n_words, op, stream ? static_cast<cudaStream_t>(stream) : g_ctx.stream);
}

int vpb_shell_dev(const uint32_t* words, uint32_t n, uint32_t* shell, void* stream) {
    VPB_TRY(require_ready());
    return shell_launch(words, n, shell, stream ? static_cast<cudaStream_t>(stream) : g_ctx.stream);
}

int vpb_csg_shell_dev(const uint32_t* a, const uint32_t* b, uint32_t n, int op, uint32_t* result, uint32_t* shell, void* stream) {
    VPB_TRY(require_ready());
    return csg_shell_launch(a, b, n, op, result, shell, stream ? static_cast<cudaStream_t>(stream) : g_ctx.stream);
}

size_t vpb_jfa_state_bytes(uint32_t n, uint32_t z0, uint32_t z1) {
    if (z1 <= z0 || z1 > n) return 0;
    return (size_t)n * n * (z1 - z0) * state_size(n);
}

int vpb_jfa_seed_dev(const uint32_t* words_full, uint32_t n, uint32_t z0, uint32_t z1, uint32_t* state, void* stream) {
    VPB_TRY(require_ready());
    return seed_any(words_full, n, z0, z1, state, stream ? static_cast<cudaStream_t>(stream) : g_ctx.stream);
}

int vpb_jfa_early_supported(uint32_t n, float vs, const float origin[3]) {
    if (!origin || n == 0 || n > kMaxJfaN) return 0;
    const Frame f = make_frame(n, vs, origin);
    return jfa_state64(n) ? jfa_early_supported_s64(f) : jfa_early_supported(f);
}

int vpb_jfa_early_dev(const uint32_t* words_full, uint32_t n, uint32_t z0, uint32_t z1, float vs, const float origin[3],
                      uint32_t* shell_scratch, uint32_t* state, void* stream) {
    VPB_TRY(require_ready());
    VPB_REQUIRE(origin && words_full && shell_scratch && state, "jfa_early: null argument");
    VPB_REQUIRE(n > 0 && n <= kMaxJfaN && z0 < z1 && z1 <= n, "jfa_early: bad n=%u slab [%u,%u)", n, z0, z1);
    return early_any(words_full, make_frame(n, vs, origin), z0, z1, shell_scratch, state,
                     stream ? static_cast<cudaStream_t>(stream) : g_ctx.stream);
}

int vpb_jfa_early_from_shell_dev(const uint32_t* shell, uint32_t n, uint32_t z0, uint32_t z1, float vs, const float origin[3],
                                 uint32_t* state, void* stream) {
    VPB_TRY(require_ready());
    VPB_REQUIRE(origin && shell && state, "jfa_early_from_shell: null argument");
    VPB_REQUIRE(n > 0 && n <= kMaxJfaN && z0 < z1 && z1 <= n, "jfa_early_from_shell: bad n=%u slab [%u,%u)", n, z0, z1);
    return early_any(nullptr, make_frame(n, vs, origin), z0, z1, const_cast<uint32_t*>(shell), state,
                     stream ? static_cast<cudaStream_t>(stream) : g_ctx.stream, true);
}

int vpb_jfa_early_dist_dev(const uint32_t* words_full, uint32_t n, float vs, const float origin[3], uint32_t rz_lo,
                           uint32_t rz_hi, uint32_t slab_planes, uint32_t* const* slab_states, uint32_t world,
                           uint32_t* shell_scratch, void* stream) {
    VPB_TRY(require_ready());
    VPB_REQUIRE(origin && words_full && shell_scratch && slab_states, "jfa_early_dist: null argument");
    VPB_REQUIRE(n > 0 && n <= kMaxJfaN, "jfa_early_dist: bad n=%u", n);
    const Frame f = make_frame(n, vs, origin);
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : g_ctx.stream;
    return jfa_state64(n) ? jfa_early_dist_launch_s64(words_full, f, rz_lo, rz_hi, slab_planes, slab_states, world, shell_scratch, st)
                          : jfa_early_dist_launch(words_full, f, rz_lo, rz_hi, slab_planes, slab_states, world, shell_scratch, st);
}

int vpb_jfa_pass_dev(const uint32_t* below, const uint32_t* mid, const uint32_t* above, uint32_t* dst, uint32_t n,
                     uint32_t z0, uint32_t z1, uint32_t k, float vs, const float origin[3], const uint32_t* words_full,
                     float* sdf, uint32_t* seeds, void* stream) {
    VPB_TRY(require_ready());
    VPB_REQUIRE(origin, "jfa_pass: null origin");
    return pass_any(below, mid, above, dst, make_frame(n, vs, origin), z0, z1, k, words_full, sdf, seeds,
                    stream ? static_cast<cudaStream_t>(stream) : g_ctx.stream);
}

int vpb_jfa_pass_part_dev(const uint32_t* src_mid, uint32_t* dst, uint32_t n, uint32_t z0, uint32_t z1, uint32_t k, float vs,
                          const float origin[3], uint32_t res_step, uint32_t res_off, void* stream) {
    VPB_TRY(require_ready());
    VPB_REQUIRE(origin && src_mid && dst, "jfa_pass_part: null argument");
    VPB_REQUIRE(n > 0 && n <= kMaxJfaN && z0 < z1 && z1 <= n && k >= 1 && k < n, "jfa_pass_part: bad n=%u slab [%u,%u) k=%u", n, z0, z1, k);
    if (jfa_state64(n)) return 1;
    return jfa_pass_flood5_launch(src_mid, dst, make_frame(n, vs, origin), z0, z1, k, nullptr, nullptr, nullptr,
                                  stream ? static_cast<cudaStream_t>(stream) : g_ctx.stream, res_step, res_off, 1, 0, 1, 0);
}

int vpb_copy_planes_dev(void* dst, size_t dst_stride, const void* src, size_t src_stride, size_t plane_bytes, size_t n_planes,
                        void* stream) {
    VPB_TRY(require_ready());
    VPB_REQUIRE(dst && src && plane_bytes > 0 && dst_stride >= plane_bytes && src_stride >= plane_bytes, "copy_planes: bad argument");
    if (n_planes == 0) return VPB_OK;
    VPB_CUDA(cudaMemcpy2DAsync(dst, dst_stride, src, src_stride, plane_bytes, n_planes, cudaMemcpyDeviceToDevice,
                               stream ? static_cast<cudaStream_t>(stream) : g_ctx.stream));
    return VPB_OK;
}

int vpb_jfa_early_cyclic_dev(const uint32_t* words_full, uint32_t n, float vs, const float origin[3], uint32_t world, uint32_t rank,
                             uint32_t* shell_scratch, uint32_t* state_cyclic, void* stream) {
    VPB_TRY(require_ready());
    VPB_REQUIRE(origin && words_full && shell_scratch && state_cyclic, "jfa_early_cyclic: null argument");
    VPB_REQUIRE(n > 0 && n <= kMaxJfaN && world >= 1 && rank < world, "jfa_early_cyclic: bad n=%u rank %u of %u", n, rank, world);
    const Frame f = make_frame(n, vs, origin);
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : g_ctx.stream;
    return jfa_state64(n) ? jfa_early_cyclic_launch_s64(words_full, f, world, rank, shell_scratch, state_cyclic, st)
                          : jfa_early_cyclic_launch(words_full, f, world, rank, shell_scratch, state_cyclic, st);
}

int vpb_jfa_pass_cyclic_dev(const uint32_t* src, uint32_t* dst, uint32_t n, uint32_t world, uint32_t rank, uint32_t k,
                            uint32_t plane_lo, uint32_t plane_hi, float vs, const float origin[3], void* stream) {
    VPB_TRY(require_ready());
    VPB_REQUIRE(origin && src && dst, "jfa_pass_cyclic: null argument");
    VPB_REQUIRE(n > 0 && n <= kMaxJfaN && world >= 1 && rank < world && n % world == 0 && k >= 1 && k < n,
                "jfa_pass_cyclic: bad n=%u rank %u of %u k=%u", n, rank, world, k);
    VPB_REQUIRE(plane_lo < plane_hi && plane_hi <= n / world, "jfa_pass_cyclic: bad plane range [%u,%u) of %u", plane_lo, plane_hi, n / world);
    if (k % world != 0) return 1;
    const Frame f = make_frame(n, vs, origin);
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : g_ctx.stream;
    if (jfa_state64(n)) return jfa_pass_flood4_cyclic_launch_s64(src, dst, f, plane_lo, plane_hi, k, world, rank, st);
    const size_t off = (size_t)plane_lo * n * n;
    const int rc = jfa_pass_flood5_launch(src + off, dst + off, f, plane_lo, plane_hi, k, nullptr, nullptr, nullptr, st, 1, 0, world, rank, 1, 0);
    return rc == 1 ? jfa_pass_flood4_cyclic_launch(src, dst, f, plane_lo, plane_hi, k, world, rank, st) : rc;
}

int vpb_jfa_pass_cyclic_to_slab_dev(const uint32_t* src, uint32_t* dst_slab, uint32_t n, uint32_t world, uint32_t rank, uint32_t k,
                                    uint32_t plane_lo, uint32_t plane_hi, float vs, const float origin[3], void* stream) {
    VPB_TRY(require_ready());
    VPB_REQUIRE(origin && src && dst_slab, "jfa_pass_cyclic_to_slab: null argument");
    VPB_REQUIRE(n > 0 && n <= kMaxJfaN && world >= 1 && rank < world && n % world == 0 && k >= 1 && k < n,
                "jfa_pass_cyclic_to_slab: bad n=%u rank %u of %u k=%u", n, rank, world, k);
    VPB_REQUIRE(plane_lo < plane_hi && plane_hi <= n / world, "jfa_pass_cyclic_to_slab: bad plane range [%u,%u) of %u", plane_lo, plane_hi, n / world);
    if (jfa_state64(n) || k % world != 0) return 1;
    const size_t off = (size_t)plane_lo * n * n;
    return jfa_pass_flood5_launch(src + off, dst_slab, make_frame(n, vs, origin), plane_lo, plane_hi, k, nullptr, nullptr, nullptr,
                                  stream ? static_cast<cudaStream_t>(stream) : g_ctx.stream, 1, 0, world, rank, world, rank);
}

int vpb_jfa_pass_peer_dev(const uint32_t* const* slab_states, uint32_t world, uint32_t slab_planes, uint32_t* dst,
                          uint32_t n, uint32_t z0, uint32_t z1, uint32_t k, float vs, const float origin[3],
                          const uint32_t* words_full, float* sdf, uint32_t* seeds, void* stream) {
    VPB_TRY(require_ready());
    VPB_REQUIRE(origin && slab_states && dst, "jfa_pass_peer: null argument");
    VPB_REQUIRE(n > 0 && n <= 1024 && z0 < z1 && z1 <= n && k >= 1 && k < n, "jfa_pass_peer: bad n=%u slab [%u,%u) k=%u", n, z0, z1, k);
    VPB_REQUIRE(!jfa_state64(n), "jfa_pass_peer: the peer-memory pass only exists for the 32-bit state");
    VPB_REQUIRE(!sdf || words_full, "jfa_pass_peer: final pass needs the occupancy grid for the sign");
    return jfa_pass_flood_peer_launch(slab_states, world, slab_planes, dst, make_frame(n, vs, origin), z0, z1, k, words_full,
                                      sdf, seeds, stream ? static_cast<cudaStream_t>(stream) : g_ctx.stream);
}

int vpb_jfa_finalize_dev(const uint32_t* state, uint32_t n, uint32_t z0, uint32_t z1, float vs, const float origin[3],
                         const uint32_t* words_full, float* sdf, uint32_t* seeds, void* stream) {
    VPB_TRY(require_ready());
    VPB_REQUIRE(origin, "jfa_finalize: null origin");
    return finalize_any(state, make_frame(n, vs, origin), z0, z1, words_full, sdf, seeds,
                        stream ? static_cast<cudaStream_t>(stream) : g_ctx.stream);
}

int vpb_jfa_dev(const uint32_t* words, uint32_t n, float vs, const float origin[3], uint32_t* sa, uint32_t* sb,
                float* sdf, uint32_t* seeds, void* stream) {
    VPB_TRY(require_ready());
    VPB_REQUIRE(words && sa && sb && sdf && origin, "jfa: null buffer");
    VPB_REQUIRE(n > 0 && n <= kMaxJfaN, "jfa: unsupported n=%u (N <= %u)", n, kMaxJfaN);
    VPB_REQUIRE(!seeds || n <= 1024, "jfa: the public 10-bit seed encoding needs N <= 1024");
    return jfa_run(words, make_frame(n, vs, origin), sa, sb, sdf, seeds, stream ? static_cast<cudaStream_t>(stream) : g_ctx.stream);
}

// ---------------------------------------------------------------------------------- host-buffer calls

static int finish_timing() {
    cudaStream_t st = g_ctx.stream;
    VPB_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < 3; ++i) {
        float ms = 0;
        VPB_CUDA(cudaEventElapsedTime(&ms, g_ctx.ev[i], g_ctx.ev[i + 1]));
        g_ctx.timing[i] = ms;
    }
    return VPB_OK;
}

// Mesh::FacesCoords of a HOST mesh must index Mesh::Coords: an OBJ with a face reference the importer could not parse (or a
// 0 / negative index) would otherwise make tri_setup read verts[3 * 0xFFFFFFFF] on the device and poison the context.
static int check_indices(const uint32_t* tris, uint64_t n_tris, uint64_t n_verts) {
    uint32_t mx = 0;
    for (uint64_t i = 0; i < 3 * n_tris; ++i) mx = tris[i] > mx ? tris[i] : mx;
    VPB_REQUIRE(n_tris == 0 || mx < n_verts, "mesh: face index %u out of range (%llu vertices)", mx, (unsigned long long)n_verts);
    return VPB_OK;
}

static int upload_mesh(const float* verts, uint64_t n_verts, const uint32_t* tris, uint64_t n_tris, cudaStream_t st) {
    VPB_TRY(check_indices(tris, n_tris, n_verts));
    VPB_TRY(g_ctx.verts.reserve(n_verts * 12 + 16));
    VPB_TRY(g_ctx.tris.reserve(n_tris * 12 + 16));
    if (n_verts) VPB_CUDA(cudaMemcpyAsync(g_ctx.verts.p, verts, n_verts * 12, cudaMemcpyHostToDevice, st));
    if (n_tris) VPB_CUDA(cudaMemcpyAsync(g_ctx.tris.p, tris, n_tris * 12, cudaMemcpyHostToDevice, st));
    return VPB_OK;
}

int vpb_voxelize_host(const float* verts, uint64_t n_verts, const uint32_t* tris, uint64_t n_tris, uint32_t n, float vs,
                      const float origin[3], int mode, uint32_t* words_out) {
    VPB_TRY(require_ready());
    VPB_REQUIRE(words_out && origin && n > 0, "voxelize: bad argument");
    VPB_REQUIRE(mode == VPB_MODE_SOLID || mode == VPB_MODE_SURFACE || mode == VPB_MODE_SURFACE_CONSERVATIVE, "voxelize: bad mode %d", mode);
    VPB_REQUIRE(n_tris == 0 || (verts && tris), "voxelize: null mesh");
    cudaStream_t st = g_ctx.stream;
    const uint64_t nw = grid_words(n);
    const size_t sb = std::max(vox_scratch_bytes(n, n_tris, 0, n), vox_surface_scratch_bytes(n_tris));
    VPB_TRY(g_ctx.grid_a.reserve(nw * 4 + 16));
    VPB_TRY(g_ctx.scratch.reserve(sb));
    if (mode == VPB_MODE_SURFACE) VPB_TRY(g_ctx.grid_b.reserve(nw * 4 + 16));
    VPB_CUDA(cudaEventRecord(g_ctx.ev[0], st));
    VPB_TRY(upload_mesh(verts, n_verts, tris, n_tris, st));
    VPB_CUDA(cudaEventRecord(g_ctx.ev[1], st));
    if (mode == VPB_MODE_SURFACE_CONSERVATIVE)
        VPB_TRY(vox_surface_launch(g_ctx.verts.as<float>(), n_verts, g_ctx.tris.as<uint32_t>(), n_tris, make_frame(n, vs, origin),
                                   0, n, g_ctx.grid_a.as<uint32_t>(), g_ctx.scratch.p, g_ctx.scratch.cap, st));
    else
        VPB_TRY(vox_launch(g_ctx.verts.as<float>(), n_verts, g_ctx.tris.as<uint32_t>(), n_tris, make_frame(n, vs, origin), 0, n,
                           g_ctx.grid_a.as<uint32_t>(), g_ctx.scratch.p, g_ctx.scratch.cap, st));
    const uint32_t* result = g_ctx.grid_a.as<uint32_t>();
    if (mode == VPB_MODE_SURFACE) {
        VPB_TRY(shell_launch(g_ctx.grid_a.as<uint32_t>(), n, g_ctx.grid_b.as<uint32_t>(), st));
        result = g_ctx.grid_b.as<uint32_t>();
    }
    VPB_CUDA(cudaEventRecord(g_ctx.ev[2], st));
    VPB_CUDA(cudaMemcpyAsync(words_out, result, nw * 4, cudaMemcpyDeviceToHost, st));
    VPB_CUDA(cudaEventRecord(g_ctx.ev[3], st));
    return finish_timing();
}

int vpb_csg_host(uint32_t* a, const uint32_t* b, uint32_t n, int op) {
    VPB_TRY(require_ready());
    VPB_REQUIRE(a && b && n > 0, "csg: bad argument");
    VPB_REQUIRE(op >= VPB_OP_UNION && op <= VPB_OP_DIFFERENCE, "csg: bad op %d", op);
    cudaStream_t st = g_ctx.stream;
    const uint64_t nw = grid_words(n);
    VPB_TRY(g_ctx.grid_a.reserve(nw * 4 + 16));
    VPB_TRY(g_ctx.grid_b.reserve(nw * 4 + 16));
    VPB_CUDA(cudaEventRecord(g_ctx.ev[0], st));
    VPB_CUDA(cudaMemcpyAsync(g_ctx.grid_a.p, a, nw * 4, cudaMemcpyHostToDevice, st));
    VPB_CUDA(cudaMemcpyAsync(g_ctx.grid_b.p, b, nw * 4, cudaMemcpyHostToDevice, st));
    VPB_CUDA(cudaEventRecord(g_ctx.ev[1], st));
    VPB_TRY(csg_launch(g_ctx.grid_a.as<uint32_t>(), g_ctx.grid_b.as<uint32_t>(), nw, op, st));
    VPB_CUDA(cudaEventRecord(g_ctx.ev[2], st));
    VPB_CUDA(cudaMemcpyAsync(a, g_ctx.grid_a.p, nw * 4, cudaMemcpyDeviceToHost, st));
    VPB_CUDA(cudaEventRecord(g_ctx.ev[3], st));
    return finish_timing();
}

// grids for which jfa_run issues the final pass in z-chunks with overlapped D2H (must match the test inside jfa_run)
static bool sink_takes(uint32_t n) { return n >= 256 && n / 2 != 0; }

// two state buffers; the signed distance is written into whichever of them the final pass leaves free (jfa_run)
static int reserve_jfa(uint32_t n, bool want_seeds) {
    const size_t vox = (size_t)n * n * n;
    VPB_TRY(g_ctx.state_a.reserve(vox * state_size(n)));
    VPB_TRY(g_ctx.state_b.reserve(vox * state_size(n)));
    if (want_seeds) VPB_TRY(g_ctx.seeds.reserve(vox * 4));
    return VPB_OK;
}

int vpb_jfa_host(const uint32_t* words, uint32_t n, float vs, const float origin[3], float* sdf_out, uint32_t* seeds_out) {
    VPB_TRY(require_ready());
    VPB_REQUIRE(words && sdf_out && origin, "jfa: null buffer");
    VPB_REQUIRE(n > 0 && n <= kMaxJfaN, "jfa: unsupported n=%u (N <= %u)", n, kMaxJfaN);
    VPB_REQUIRE(!seeds_out || n <= 1024, "jfa: the public 10-bit seed encoding needs N <= 1024");
    cudaStream_t st = g_ctx.stream;
    const uint64_t nw = grid_words(n);
    const size_t vox = (size_t)n * n * n;
    VPB_TRY(g_ctx.grid_a.reserve(nw * 4 + 16));
    VPB_TRY(reserve_jfa(n, seeds_out != nullptr));
    VPB_CUDA(cudaEventRecord(g_ctx.ev[0], st));
    VPB_CUDA(cudaMemcpyAsync(g_ctx.grid_a.p, words, nw * 4, cudaMemcpyHostToDevice, st));
    VPB_CUDA(cudaEventRecord(g_ctx.ev[1], st));
    float* sdf_dev = nullptr;
    const bool chunked = sink_takes(n);
    HostSink sink;
    sink.sdf_host = sdf_out; sink.seeds_host = seeds_out; sink.kernels_done = g_ctx.ev[2];
    VPB_TRY(jfa_run(g_ctx.grid_a.as<uint32_t>(), make_frame(n, vs, origin), g_ctx.state_a.as<uint32_t>(),
                    g_ctx.state_b.as<uint32_t>(), nullptr, seeds_out ? g_ctx.seeds.as<uint32_t>() : nullptr, st, &sdf_dev,
                    chunked ? &sink : nullptr));
    if (!chunked) {
        VPB_CUDA(cudaEventRecord(g_ctx.ev[2], st));
        VPB_CUDA(cudaMemcpyAsync(sdf_out, sdf_dev, vox * 4, cudaMemcpyDeviceToHost, st));
        if (seeds_out) VPB_CUDA(cudaMemcpyAsync(seeds_out, g_ctx.seeds.p, vox * 4, cudaMemcpyDeviceToHost, st));
    }
    VPB_CUDA(cudaEventRecord(g_ctx.ev[3], st));
    return finish_timing();
}

// Voxelizes every mesh in the shared frame and folds grids[0] = op(grids[0], grids[i]) (apps/cli/main.cpp:92-186); the
// folded grid ends up in `final_dst`.  With `shell_dst`, the LAST fold also writes the seed shell of the result there
// (csg_shell_launch: CSG fused with the seed extraction) and *shell_ready tells; `acc` is then the accumulator of the fold
// (it must differ from final_dst), otherwise everything happens in final_dst.
static int occupancy_stage(int n_meshes, const float* const* verts, const uint64_t* n_verts, const uint32_t* const* tris,
                           const uint64_t* n_tris, const Frame& f, int op, uint32_t* final_dst, uint32_t* acc, uint32_t* operand,
                           uint32_t* shell_dst, bool* shell_ready, cudaEvent_t first_upload_done, cudaStream_t st) {
    const uint64_t nw = grid_words(f.n);
    const bool fuse = shell_dst && acc && n_meshes >= 2 && op != VPB_OP_VOID && f.n % 32u == 0 && early_takes(f);
    *shell_ready = false;
    uint32_t* a = fuse ? acc : final_dst;
    for (int i = 0; i < n_meshes; ++i) {
        VPB_TRY(upload_mesh(verts[i], n_verts[i], tris[i], n_tris[i], st));
        if (i == 0 && first_upload_done) VPB_CUDA(cudaEventRecord(first_upload_done, st));
        uint32_t* target = i == 0 ? a : operand;
        VPB_TRY(vox_launch(g_ctx.verts.as<float>(), n_verts[i], g_ctx.tris.as<uint32_t>(), n_tris[i], f, 0, f.n, target,
                           g_ctx.scratch.p, g_ctx.scratch.cap, st));
        if (i == 0 || op == VPB_OP_VOID) continue;
        if (fuse && i == n_meshes - 1) {
            const int rc = csg_shell_launch(a, operand, f.n, op, final_dst, shell_dst, st);
            if (rc < 0) return rc;
            *shell_ready = rc == 0;
            if (rc != 0) {                                   // not taken after all: fold in place and move the result
                VPB_TRY(csg_launch(a, operand, nw, op, st));
                VPB_CUDA(cudaMemcpyAsync(final_dst, a, nw * 4, cudaMemcpyDeviceToDevice, st));
            }
        } else {
            VPB_TRY(csg_launch(a, operand, nw, op, st));
        }
    }
    return VPB_OK;
}

int vpb_pipeline_host(int n_meshes, const float* const* verts, const uint64_t* n_verts, const uint32_t* const* tris,
                      const uint64_t* n_tris, uint32_t n, float vs, const float origin[3], int op, uint32_t* words_out,
                      float* sdf_out) {
    VPB_TRY(require_ready());
    VPB_REQUIRE(n_meshes >= 1 && verts && n_verts && tris && n_tris && origin && n > 0, "pipeline: bad argument");
    VPB_REQUIRE(op >= VPB_OP_VOID && op <= VPB_OP_DIFFERENCE, "pipeline: bad op %d", op);
    VPB_REQUIRE(!sdf_out || n <= kMaxJfaN, "pipeline: sdf needs N <= %u", kMaxJfaN);
    cudaStream_t st = g_ctx.stream;
    const uint64_t nw = grid_words(n);
    const size_t vox = (size_t)n * n * n;
    const Frame f = make_frame(n, vs, origin);
    uint64_t max_tris = 0;
    for (int i = 0; i < n_meshes; ++i) max_tris = n_tris[i] > max_tris ? n_tris[i] : max_tris;
    VPB_TRY(g_ctx.grid_a.reserve(nw * 4 + 16));
    if (n_meshes > 1) VPB_TRY(g_ctx.grid_b.reserve(nw * 4 + 16));
    VPB_TRY(g_ctx.scratch.reserve(vox_scratch_bytes(n, max_tris, 0, n)));
    if (sdf_out) VPB_TRY(reserve_jfa(n, false));
    // H2D of mesh i and its kernels are interleaved on one stream; ev[0..1] bracket the first upload only,
    // the rest is accounted to "kernels" (there is a single timeline).
    const bool want_fuse = sdf_out && n_meshes >= 2 && op != VPB_OP_VOID && n % 32u == 0;
    if (want_fuse) VPB_TRY(g_ctx.grid_c.reserve(nw * 4 + 16));
    VPB_CUDA(cudaEventRecord(g_ctx.ev[0], st));
    bool shell_ready = false;
    VPB_TRY(occupancy_stage(n_meshes, verts, n_verts, tris, n_tris, f, op, g_ctx.grid_a.as<uint32_t>(),
                            want_fuse ? g_ctx.grid_c.as<uint32_t>() : nullptr, g_ctx.grid_b.as<uint32_t>(),
                            want_fuse ? g_ctx.state_b.as<uint32_t>() : nullptr, &shell_ready, g_ctx.ev[1], st));
    float* sdf_dev = nullptr;
    const bool chunked = sdf_out && sink_takes(n);
    if (sdf_out) {
        HostSink sink;
        sink.sdf_host = sdf_out; sink.kernels_done = g_ctx.ev[2];
        VPB_TRY(jfa_run(g_ctx.grid_a.as<uint32_t>(), f, g_ctx.state_a.as<uint32_t>(), g_ctx.state_b.as<uint32_t>(),
                        nullptr, nullptr, st, &sdf_dev, chunked ? &sink : nullptr, shell_ready));
    }
    if (!chunked) VPB_CUDA(cudaEventRecord(g_ctx.ev[2], st));
    if (words_out) VPB_CUDA(cudaMemcpyAsync(words_out, g_ctx.grid_a.p, nw * 4, cudaMemcpyDeviceToHost, st));
    if (sdf_out && !chunked) VPB_CUDA(cudaMemcpyAsync(sdf_out, sdf_dev, vox * 4, cudaMemcpyDeviceToHost, st));
    VPB_CUDA(cudaEventRecord(g_ctx.ev[3], st));
    return finish_timing();
}

// ---- asynchronous form of vpb_pipeline_host: up to two jobs in flight ------------------------------------------------
// Job j's kernels run on the compute stream; every finished z-chunk of its sdf (and its occupancy words) is copied to the
// host on the copy stream from buffers only job j+2 will reuse, so job j+1's kernels overlap job j's D2H -- the host
// API is PCIe-bound (4.3 GB of sdf per 1024^3 job), not kernel-bound.
int vpb_pipeline_submit(int n_meshes, const float* const* verts, const uint64_t* n_verts, const uint32_t* const* tris,
                        const uint64_t* n_tris, uint32_t n, float vs, const float origin[3], int op, uint32_t* words_out,
                        float* sdf_out, uint64_t* ticket) {
    VPB_TRY(require_ready());
    VPB_REQUIRE(n_meshes >= 1 && verts && n_verts && tris && n_tris && origin && n > 0 && ticket, "pipeline_submit: bad argument");
    VPB_REQUIRE(op >= VPB_OP_VOID && op <= VPB_OP_DIFFERENCE, "pipeline_submit: bad op %d", op);
    VPB_REQUIRE(!sdf_out || n <= 1024, "pipeline_submit: sdf needs N <= 1024 (two result slots must fit the device)");
    cudaStream_t st = g_ctx.stream, cs = g_ctx.copy_stream;
    const uint64_t nw = grid_words(n);
    const size_t vox = (size_t)n * n * n;
    const Frame f = make_frame(n, vs, origin);
    const uint64_t t = g_ctx.next_ticket;
    const int slot = (int)(t & 1u);
    // the slot's previous job (ticket t-2) must have left the device before its buffers are overwritten
    if (g_ctx.slot_ticket[slot]) VPB_CUDA(cudaStreamWaitEvent(st, g_ctx.slot_done[slot], 0));
    uint64_t max_tris = 0;
    for (int i = 0; i < n_meshes; ++i) max_tris = n_tris[i] > max_tris ? n_tris[i] : max_tris;
    VPB_TRY(g_ctx.grid_a.reserve(nw * 4 + 16));
    if (n_meshes > 1) VPB_TRY(g_ctx.grid_b.reserve(nw * 4 + 16));
    VPB_TRY(g_ctx.scratch.reserve(vox_scratch_bytes(n, max_tris, 0, n)));
    VPB_TRY(g_ctx.slot_words[slot].reserve(nw * 4 + 16));
    if (sdf_out) {
        VPB_TRY(reserve_jfa(n, false));
        VPB_TRY(g_ctx.slot_sdf[slot].reserve(vox * 4));
    }
    uint32_t* words = g_ctx.slot_words[slot].as<uint32_t>();          // grids[0] of this job lives in the slot
    bool shell_ready = false;
    VPB_TRY(occupancy_stage(n_meshes, verts, n_verts, tris, n_tris, f, op, words, g_ctx.grid_a.as<uint32_t>(),
                            g_ctx.grid_b.as<uint32_t>(), sdf_out ? g_ctx.state_b.as<uint32_t>() : nullptr, &shell_ready, nullptr, st));
    VPB_CUDA(cudaEventRecord(g_ctx.chunk_ev[15], st));                 // occupancy final
    bool copies_recorded = false;
    if (sdf_out) {
        HostSink sink;
        sink.sdf_host = sdf_out;
        sink.copies_done = g_ctx.slot_done[slot];
        const bool chunked = sink_takes(n);
        if (words_out && chunked) {                                    // ahead of the sdf chunks on the copy stream
            VPB_CUDA(cudaStreamWaitEvent(cs, g_ctx.chunk_ev[15], 0));
            VPB_CUDA(cudaMemcpyAsync(words_out, words, nw * 4, cudaMemcpyDeviceToHost, cs));
        }
        VPB_TRY(jfa_run(words, f, g_ctx.state_a.as<uint32_t>(), g_ctx.state_b.as<uint32_t>(), g_ctx.slot_sdf[slot].as<float>(),
                        nullptr, st, nullptr, chunked ? &sink : nullptr, shell_ready));
        copies_recorded = chunked;
        if (!chunked) {
            VPB_CUDA(cudaEventRecord(g_ctx.chunk_ev[14], st));
            VPB_CUDA(cudaStreamWaitEvent(cs, g_ctx.chunk_ev[14], 0));
            if (words_out) VPB_CUDA(cudaMemcpyAsync(words_out, words, nw * 4, cudaMemcpyDeviceToHost, cs));
            VPB_CUDA(cudaMemcpyAsync(sdf_out, g_ctx.slot_sdf[slot].p, vox * 4, cudaMemcpyDeviceToHost, cs));
        }
    } else if (words_out) {
        VPB_CUDA(cudaStreamWaitEvent(cs, g_ctx.chunk_ev[15], 0));
        VPB_CUDA(cudaMemcpyAsync(words_out, words, nw * 4, cudaMemcpyDeviceToHost, cs));
    }
    if (!copies_recorded) {
        VPB_CUDA(cudaStreamWaitEvent(cs, g_ctx.chunk_ev[15], 0));      // also covers "no output requested"
        VPB_CUDA(cudaEventRecord(g_ctx.slot_done[slot], cs));
    }
    g_ctx.slot_ticket[slot] = t;
    g_ctx.next_ticket = t + 1;
    *ticket = t;
    return VPB_OK;
}

int vpb_pipeline_wait(uint64_t ticket) {
    VPB_TRY(require_ready());
    const int slot = (int)(ticket & 1u);
    VPB_REQUIRE(ticket != 0 && ticket < g_ctx.next_ticket, "pipeline_wait: unknown ticket %llu", (unsigned long long)ticket);
    // (if the slot has been reused by ticket + 2 the event now marks that later job's copies: waiting for them is
    // sufficient, the copy stream is in order)
    VPB_CUDA(cudaEventSynchronize(g_ctx.slot_done[slot]));
    // kernel errors of the job surface here
    VPB_CUDA(cudaGetLastError());
    return VPB_OK;
}

}  // extern "C"
