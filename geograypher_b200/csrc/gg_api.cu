// extern "C" surface of libgeograypher_b200.so (declared in include/geograypher_b200.h).
#include <cstdio>
#include <cstring>

#include <cstdlib>
#include <algorithm>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <unistd.h>
#include <vector>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_reduce.cuh>

#include "gg_internal.cuh"

static thread_local std::string g_last_error;

void gg_set_error(const std::string &msg) { g_last_error = msg; }

int gg_cuda_fail(cudaError_t e, const char *what) {
    g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
    return GG_ERR_CUDA;
}

namespace {
__global__ void k_pack_mesh(const float *__restrict__ v, int64_t V, const int32_t *__restrict__ f, int64_t F,
                            float4 *__restrict__ v4, int4 *__restrict__ f4, int *__restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < V) v4[i] = make_float4(v[3 * i], v[3 * i + 1], v[3 * i + 2], 0.f);
    if (i < F) {
        const int a = f[3 * i], b = f[3 * i + 1], c = f[3 * i + 2];
        if (a < 0 || b < 0 || c < 0 || a >= V || b >= V || c >= V) atomicOr(bad, 1);
        f4[i] = make_int4(a, b, c, (int)i);
    }
}
// ---- spatial ordering of the faces (Morton order of the centroids) -----------------------------------------
// The frustum culling works on blocks of 128 consecutive faces; sorting the faces along a Z-order curve makes those
// blocks compact whatever order the mesh file had.  Face IDs travel with the faces (int4.w), so results are unchanged.
__device__ __forceinline__ unsigned spread_bits10(unsigned v) {
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

__global__ void k_vertex_bounds_init(float *lohi) {
    if (threadIdx.x < 3) {
        lohi[threadIdx.x] = INFINITY;
        lohi[3 + threadIdx.x] = -INFINITY;
    }
}

__device__ __forceinline__ void atomic_min_float(float *addr, float v) {  // valid for any sign
    if (v >= 0) atomicMin(reinterpret_cast<int *>(addr), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned *>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_float(float *addr, float v) {
    if (v >= 0) atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned *>(addr), __float_as_uint(v));
}

__global__ void __launch_bounds__(256) k_vertex_bounds(const float4 *__restrict__ v4, int64_t V, float *lohi) {
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < V; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = v4[i];
        if (isfinite(v.x) && isfinite(v.y) && isfinite(v.z)) {
            mn[0] = fminf(mn[0], v.x); mx[0] = fmaxf(mx[0], v.x);
            mn[1] = fminf(mn[1], v.y); mx[1] = fmaxf(mx[1], v.y);
            mn[2] = fminf(mn[2], v.z); mx[2] = fmaxf(mx[2], v.z);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
        if ((threadIdx.x & 31) == 0) {
            if (mn[k] < INFINITY) atomic_min_float(&lohi[k], mn[k]);
            if (mx[k] > -INFINITY) atomic_max_float(&lohi[3 + k], mx[k]);
        }
    }
}

__global__ void __launch_bounds__(256) k_morton_codes(const float4 *__restrict__ v4, const int4 *__restrict__ f4, int64_t F,
                                                      const float *__restrict__ lohi, unsigned *__restrict__ codes,
                                                      int *__restrict__ order) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F) return;
    const int4 f = f4[i];
    const float4 a = v4[f.x], b = v4[f.y], c = v4[f.z];
    const float cen[3] = {(a.x + b.x + c.x) * (1.f / 3.f), (a.y + b.y + c.y) * (1.f / 3.f), (a.z + b.z + c.z) * (1.f / 3.f)};
    unsigned code = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float ext = lohi[3 + k] - lohi[k];
        float t = ext > 0.f ? (cen[k] - lohi[k]) / ext : 0.f;
        t = isfinite(t) ? fminf(fmaxf(t, 0.f), 1.f) : 0.f;
        code |= spread_bits10((unsigned)(t * 1023.f)) << k;
    }
    codes[i] = code;
    order[i] = (int)i;
}

__global__ void __launch_bounds__(256) k_permute_faces(const int4 *__restrict__ in, const int *__restrict__ order, int64_t F,
                                                       int4 *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < F) out[i] = in[order[i]];
}
}  // namespace

static int sort_faces_spatially(gg_context *ctx, cudaStream_t st) {
    const int64_t F = ctx->F, V = ctx->V;
    if (F < 2 * GG_BLOCK_FACES || V == 0) return GG_OK;  // a single cull block: nothing to gain
    // one allocation for all temporaries, so that every exit path frees exactly one pointer
    const size_t nF = (size_t)F;
    const size_t off_codes = 256, off_codes_out = off_codes + nF * 4, off_order = off_codes_out + nF * 4,
                 off_order_out = off_order + nF * 4;
    size_t tmp_bytes = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const unsigned *)nullptr, (unsigned *)nullptr,
                                                    (const int *)nullptr, (int *)nullptr, (int)F, 0, 30, st);
    if (e != cudaSuccess) return gg_cuda_fail(e, "cub::DeviceRadixSort::SortPairs (size query)");
    const size_t off_tmp = (off_order_out + nF * 4 + 255) / 256 * 256;
    char *buf = nullptr;
    int4 *d_sorted = nullptr;
    GG_CUDA(cudaMalloc(&buf, off_tmp + tmp_bytes));
    e = cudaMalloc(&d_sorted, nF * sizeof(int4));
    if (e == cudaSuccess) {
        float *d_lohi = (float *)buf;
        unsigned *d_codes = (unsigned *)(buf + off_codes), *d_codes_out = (unsigned *)(buf + off_codes_out);
        int *d_order = (int *)(buf + off_order), *d_order_out = (int *)(buf + off_order_out);
        k_vertex_bounds_init<<<1, 32, 0, st>>>(d_lohi);
        k_vertex_bounds<<<ctx->sm_count * 4, 256, 0, st>>>(ctx->d_verts, V, d_lohi);
        k_morton_codes<<<(unsigned)((F + 255) / 256), 256, 0, st>>>(ctx->d_verts, ctx->d_faces, F, d_lohi, d_codes, d_order);
        e = cudaGetLastError();
        if (e == cudaSuccess)
            e = cub::DeviceRadixSort::SortPairs(buf + off_tmp, tmp_bytes, d_codes, d_codes_out, d_order, d_order_out, (int)F, 0,
                                                30, st);
        if (e == cudaSuccess) {
            k_permute_faces<<<(unsigned)((F + 255) / 256), 256, 0, st>>>(ctx->d_faces, d_order_out, F, d_sorted);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    cudaFree(buf);
    if (e != cudaSuccess) {
        cudaFree(d_sorted);
        return gg_cuda_fail(e, "sorting the faces along the Z-order curve");
    }
    cudaFree(ctx->d_faces);
    ctx->d_faces = d_sorted;
    return GG_OK;
}

static int check_ctx(gg_context *ctx, bool need_mesh) {
    if (!ctx) {
        gg_set_error("null context");
        return GG_ERR_INVALID;
    }
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return gg_cuda_fail(e, "cudaSetDevice");
    if (need_mesh && ctx->F == 0) {
        gg_set_error("gg_set_mesh has not been called");
        return GG_ERR_NO_MESH;
    }
    return GG_OK;
}

static int check_cams(const gg_camera *cams, int n) {
    if (!cams || n < 1 || n > GG_MAX_VIEWS_PER_CALL) {
        gg_set_error("need 1..GG_MAX_VIEWS_PER_CALL cameras");
        return GG_ERR_INVALID;
    }
    for (int i = 0; i < n; ++i) {
        if (cams[i].W != cams[0].W || cams[i].H != cams[0].H) {
            gg_set_error("Not all cameras have the same image size");
            return GG_ERR_INVALID;
        }
        if (cams[i].W < 1 || cams[i].H < 1 || cams[i].W > 32767 || cams[i].H > 32767) {
            gg_set_error("raster size must be within 1..32767");
            return GG_ERR_INVALID;
        }
    }
    return GG_OK;
}

static const char *k_stage_names[GG_ST_COUNT] = {"mesh_setup", "project", "cull_blocks", "setup_faces", "scan_tiles",
                                                   "fill_bins", "raster_tiles", "last_pixel", "resolve", "pixel_sum",
                                                   "finalize", "render_flat", "misc", "stage_host_rows"};

const char *gg_stage_label(int stage) { return (stage >= 0 && stage < GG_ST_COUNT) ? k_stage_names[stage] : "?"; }

extern "C" {

int gg_abi_version(void) { return GG_ABI_VERSION; }

int gg_host_alloc(int device, size_t bytes, void **out) {
    if (!out || bytes == 0) {
        gg_set_error("gg_host_alloc: bad arguments");
        return GG_ERR_INVALID;
    }
    *out = nullptr;
    GG_CUDA(cudaSetDevice(device));
    void *p = nullptr;
    GG_CUDA(cudaMallocManaged(&p, bytes, cudaMemAttachGlobal));
    // the pages live in HOST memory and are mapped into the GPU's address space: kernels read them over PCIe in
    // place, nothing migrates
    cudaError_t e = cudaMemAdvise(p, bytes, cudaMemAdviseSetPreferredLocation, cudaCpuDeviceId);
    if (e == cudaSuccess) e = cudaMemAdvise(p, bytes, cudaMemAdviseSetAccessedBy, device);
    if (e != cudaSuccess) {
        cudaFree(p);
        return gg_cuda_fail(e, "cudaMemAdvise");
    }
    *out = p;
    return GG_OK;
}

int gg_host_free(void *p) {
    if (p) GG_CUDA(cudaFree(p));
    return GG_OK;
}

int gg_pointer_kind(const void *p) {
    cudaPointerAttributes attr;
    if (!p || cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    switch (attr.type) {
        case cudaMemoryTypeHost: return 1;
        case cudaMemoryTypeDevice: return 2;
        case cudaMemoryTypeManaged: return 3;
        default: return 0;
    }
}

namespace {
// Persistent host threads for gg_gather_rows_host: starting 15 threads per call costs as much as the gather itself.
struct HostPool {
    std::vector<std::thread> threads;
    std::mutex m;
    std::condition_variable wake, done;
    const std::function<void(int)> *job = nullptr;
    int n_active = 0;   // workers taking part in the current job (worker t runs slice t + 1)
    int generation = 0, pending = 0;
    bool stop = false;

    void worker(int t) {
        int seen = 0;
        for (;;) {
            const std::function<void(int)> *f;
            {
                std::unique_lock<std::mutex> lk(m);
                wake.wait(lk, [&] { return stop || (generation != seen && t < n_active); });
                if (stop) return;
                seen = generation;
                f = job;
            }
            (*f)(t + 1);
            {
                std::lock_guard<std::mutex> lk(m);
                if (--pending == 0) done.notify_one();
            }
        }
    }
    void run(int T, const std::function<void(int)> &f) {  // slices 0 .. T-1; slice 0 on the calling thread
        {
            std::lock_guard<std::mutex> lk(m);
            while ((int)threads.size() < T - 1) {
                const int t = (int)threads.size();
                threads.emplace_back([this, t] { worker(t); });
            }
            job = &f;
            n_active = T - 1;
            pending = T - 1;
            ++generation;
        }
        wake.notify_all();
        f(0);
        std::unique_lock<std::mutex> lk(m);
        done.wait(lk, [&] { return pending == 0; });
        n_active = 0;
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> lk(m);
            stop = true;
        }
        wake.notify_all();
        for (auto &t : threads) t.join();
    }
};
HostPool &host_pool() {
    // never destroyed (worker threads must not be joined from atexit handlers); a forked child starts its own, since
    // threads do not survive fork()
    static HostPool *pool = nullptr;
    static pid_t owner = 0;
    if (!pool || owner != getpid()) {
        pool = new HostPool();
        owner = getpid();
    }
    return *pool;
}
std::mutex g_gather_mutex;
}  // namespace

int gg_gather_rows_host(const void *const *h_images, const int64_t *h_pixels_per_image, const int32_t *h_pairs,
                        const int64_t *h_pair_starts, const int64_t *h_offsets, int n_views, int64_t row_bytes,
                        void *h_out, int n_threads) {
    if (!h_images || !h_pixels_per_image || !h_pairs || !h_offsets || !h_out || n_views < 1 || row_bytes < 1) {
        gg_set_error("gg_gather_rows_host: bad arguments");
        return GG_ERR_INVALID;
    }
    const int64_t total = h_offsets[n_views];
    if (total <= 0) return GG_OK;
    // default: up to 16 threads (more only add wake-up latency to a gather that lasts a fraction of a millisecond)
    int T = n_threads > 0 ? n_threads : std::min(16, (int)std::thread::hardware_concurrency());
    T = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(T, 64), total / 4096 + 1));
    // pair i of view v sits at h_pairs[2 * (start[v] + i - h_offsets[v])]: packed lists by default
    const int64_t *start = h_pair_starts ? h_pair_starts : h_offsets;
    const std::function<void(int)> work = [&](int t) {
        const int64_t lo = total * t / T, hi = total * (t + 1) / T;
        int v = 0;
        while (h_offsets[v + 1] <= lo) ++v;
        int va = v;
        constexpr int kAhead = 16;  // rows are cache (and TLB) misses: keep some in flight per thread
        for (int64_t i = lo; i < hi; ++i) {
            while (h_offsets[v + 1] <= i) ++v;
            if (i + kAhead < hi) {
                while (h_offsets[va + 1] <= i + kAhead) ++va;
                int64_t pa = h_pairs[2 * (start[va] + (i + kAhead - h_offsets[va])) + 1];
                pa = pa < 0 ? 0 : (pa >= h_pixels_per_image[va] ? h_pixels_per_image[va] - 1 : pa);
                const char *a = (const char *)h_images[va] + pa * row_bytes;
                __builtin_prefetch(a);
                __builtin_prefetch(a + row_bytes - 1);
            }
            int64_t p = h_pairs[2 * (start[v] + (i - h_offsets[v])) + 1];
            p = p < 0 ? 0 : (p >= h_pixels_per_image[v] ? h_pixels_per_image[v] - 1 : p);  // np.take(mode="clip")
            memcpy((char *)h_out + i * row_bytes, (const char *)h_images[v] + p * row_bytes, (size_t)row_bytes);
        }
    };
    if (T == 1) {
        work(0);
        return GG_OK;
    }
    std::lock_guard<std::mutex> one_at_a_time(g_gather_mutex);
    host_pool().run(T, work);
    return GG_OK;
}

const char *gg_last_error(void) { return g_last_error.c_str(); }

int gg_create(int device, gg_context **out) {
    if (!out) {
        gg_set_error("gg_create: out is null");
        return GG_ERR_INVALID;
    }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0 || device < 0 || device >= count) {
        gg_set_error(std::string("gg_create: no usable CUDA device (") +
                     (e != cudaSuccess ? cudaGetErrorString(e) : "index out of range") + "); there is no CPU fallback");
        return GG_ERR_NO_DEVICE;
    }
    GG_CUDA(cudaSetDevice(device));
    gg_context *ctx = new gg_context();
    ctx->device = device;
    cudaDeviceProp prop;
    GG_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    memset(ctx->vset, 0, sizeof(ctx->vset));
    if (const char *e = getenv("GG_DENSE_PREFETCH")) ctx->dense_prefetch = atoi(e) != 0;
    if (const char *e = getenv("GG_LANE_LISTS")) {
        ctx->lane_lists = atoi(e) != 0;
        ctx->lane_lists_auto = 0;
    }
    if (const char *e = getenv("GG_LANE_LISTS_MIN_ENTRIES")) ctx->lane_lists_min_entries = (float)atof(e);
    if (const char *e = getenv("GG_STAGE_HOST_ROWS")) ctx->stage_host_rows = atoi(e) != 0;
    if (const char *e = getenv("GG_SETUP_CTAS")) ctx->setup_ctas = atoi(e);
    if (const char *e = getenv("GG_FILL_CTAS")) ctx->fill_ctas = atoi(e);
    if (const char *e = getenv("GG_STAGE_INFLIGHT")) ctx->stage_inflight = atoi(e) > 0 ? atoi(e) : ctx->stage_inflight;
    if (const char *e = getenv("GG_STAGE_GRID")) ctx->stage_grid = atoi(e);
    if (const char *e = getenv("GG_STAGE_CTAS")) ctx->stage_ctas = atoi(e) > 0 ? atoi(e) : 0;
    GG_CUDA(cudaMalloc(&ctx->d_sticky, 4 * sizeof(int32_t)));
    GG_CUDA(cudaMemset(ctx->d_sticky, 0, 4 * sizeof(int32_t)));
    // the short, latency-bound binning kernels get the higher priority so that they slip in between the CTAs of the
    // rasterizer that is still running for the previous batch
    int prio_lo = 0, prio_hi = 0;
    GG_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    GG_CUDA(cudaStreamCreateWithPriority(&ctx->sA, cudaStreamNonBlocking, prio_hi));
    GG_CUDA(cudaStreamCreateWithPriority(&ctx->sB, cudaStreamNonBlocking, prio_lo));
    GG_CUDA(cudaStreamCreateWithPriority(&ctx->sC, cudaStreamNonBlocking, prio_hi));
    GG_CUDA(cudaEventCreateWithFlags(&ctx->ev_user, cudaEventDisableTiming));
    for (int s = 0; s < GG_NSETS; ++s) {
        GG_CUDA(cudaEventCreateWithFlags(&ctx->ev_bin[s], cudaEventDisableTiming));
        GG_CUDA(cudaEventCreateWithFlags(&ctx->ev_ras[s], cudaEventDisableTiming));
        GG_CUDA(cudaEventCreateWithFlags(&ctx->ev_mid[s], cudaEventDisableTiming));
    }
    *out = ctx;
    return GG_OK;
}

void gg_destroy(gg_context *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    cudaFree(ctx->d_verts);
    cudaFree(ctx->d_faces);
    cudaFree(ctx->d_block_lo);
    cudaFree(ctx->d_block_hi);
    cudaFree(ctx->d_scratch);
    cudaFree(ctx->d_winner);
    cudaFree(ctx->d_wdense);
    cudaFree(ctx->d_stage);
    cudaFree(ctx->d_sticky);
    cudaFree(ctx->d_raster);
    for (auto &p : ctx->prof.pending) {
        cudaEventDestroy(p.a);
        cudaEventDestroy(p.b);
    }
    for (auto e : ctx->prof.pool) cudaEventDestroy(e);
    if (ctx->sA) cudaStreamDestroy(ctx->sA);
    if (ctx->sB) cudaStreamDestroy(ctx->sB);
    if (ctx->sC) cudaStreamDestroy(ctx->sC);
    if (ctx->ev_user) cudaEventDestroy(ctx->ev_user);
    for (int s = 0; s < GG_NSETS; ++s) {
        if (ctx->ev_bin[s]) cudaEventDestroy(ctx->ev_bin[s]);
        if (ctx->ev_ras[s]) cudaEventDestroy(ctx->ev_ras[s]);
        if (ctx->ev_mid[s]) cudaEventDestroy(ctx->ev_mid[s]);
    }
    delete ctx;
}

int gg_sync(gg_context *ctx, void *stream) {
    int rc = check_ctx(ctx, false);
    if (rc != GG_OK) return rc;
    GG_CUDA(cudaStreamSynchronize(ctx->sA));
    GG_CUDA(cudaStreamSynchronize(ctx->sB));
    GG_CUDA(cudaStreamSynchronize(ctx->sC));
    for (int s = 0; s < GG_NSETS; ++s) ctx->ras_pending[s] = false;
    GG_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    GG_CUDA(cudaGetLastError());
    int32_t st4[4] = {0, 0, 0, 0};
    GG_CUDA(cudaMemcpy(st4, ctx->d_sticky, sizeof(st4), cudaMemcpyDeviceToHost));
    if (st4[1] != 0 || st4[2] != 0 || st4[0] != 0) GG_CUDA(cudaMemset(ctx->d_sticky, 0, 4 * sizeof(int32_t)));
    // st4[2]: the most tile entries any view needed since the last gg_sync -> which rasterizer variant comes next
    if (ctx->lane_lists_auto && st4[2] > 0 && ctx->last_n_tiles > 0)
        ctx->lane_lists = (double)st4[2] >= (double)ctx->lane_lists_min_entries * (double)ctx->last_n_tiles;
    const int32_t sticky = st4[0];
    if (sticky != 0) {
        ctx->last_overflow[0] = sticky;
        ctx->last_overflow[1] = st4[1];
        ctx->last_overflow[2] = st4[2];
        char buf[320];
        snprintf(buf, sizeof(buf),
                 "scratch overflow since the last gg_sync (%s%s; capacity: %lld face records, %lld tile entries per "
                 "view): the affected batches were skipped; call gg_reserve with larger capacities and redo them",
                 (sticky & 1) ? "face records " : "", (sticky & 2) ? "tile entries" : "", (long long)ctx->cap_recs,
                 (long long)ctx->cap_bins);
        gg_set_error(buf);
        return GG_ERR_OVERFLOW;
    }
    return GG_OK;
}



int gg_stage_count(void) { return GG_ST_COUNT; }

const char *gg_stage_name(int stage) { return (stage >= 0 && stage < GG_ST_COUNT) ? k_stage_names[stage] : ""; }

int gg_profile(gg_context *ctx, int enable) {
    int rc = check_ctx(ctx, false);
    if (rc != GG_OK) return rc;
    ctx->prof.on = enable != 0;
    return GG_OK;
}

int gg_profile_read(gg_context *ctx, double *h_ms, int64_t *h_launches, int reset) {
    int rc = check_ctx(ctx, false);
    if (rc != GG_OK) return rc;
    GG_CUDA(cudaDeviceSynchronize());
    for (auto &p : ctx->prof.pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) ctx->prof.ms[p.stage] += ms;
        ctx->prof.pool.push_back(p.a);
        ctx->prof.pool.push_back(p.b);
    }
    ctx->prof.pending.clear();
    for (int i = 0; i < GG_ST_COUNT; ++i) {
        if (h_ms) h_ms[i] = ctx->prof.ms[i];
        if (h_launches) h_launches[i] = ctx->prof.launches[i];
        if (reset) {
            ctx->prof.ms[i] = 0;
            ctx->prof.launches[i] = 0;
        }
    }
    return GG_OK;
}

int gg_drain(gg_context *ctx, void *stream) {
    int rc = check_ctx(ctx, false);
    if (rc != GG_OK) return rc;
    return gg_pipeline_drain(ctx, (cudaStream_t)stream);
}

int gg_set_pipeline(gg_context *ctx, int enable) {
    int rc = check_ctx(ctx, false);
    if (rc != GG_OK) return rc;
    GG_CUDA(cudaDeviceSynchronize());
    for (int s = 0; s < GG_NSETS; ++s) ctx->ras_pending[s] = false;
    ctx->pipeline = enable != 0;
    return GG_OK;
}

int gg_reserve(gg_context *ctx, int64_t max_faces_per_view, int64_t max_bin_entries_per_view) {
    int rc = check_ctx(ctx, false);
    if (rc != GG_OK) return rc;
    if (max_faces_per_view < 0 || max_bin_entries_per_view < 0) {
        gg_set_error("gg_reserve: negative capacity");
        return GG_ERR_INVALID;
    }
    ctx->req_recs = max_faces_per_view;
    ctx->req_bins = max_bin_entries_per_view;
    return GG_OK;
}

int gg_get_capacity(gg_context *ctx, int64_t *h_faces_per_view, int64_t *h_bin_entries_per_view) {
    int rc = check_ctx(ctx, false);
    if (rc != GG_OK) return rc;
    if (h_faces_per_view) *h_faces_per_view = ctx->cap_recs;
    if (h_bin_entries_per_view) *h_bin_entries_per_view = ctx->cap_bins;
    return GG_OK;
}

int gg_overflow_info(gg_context *ctx, int64_t *h_out3) {
    int rc = check_ctx(ctx, false);
    if (rc != GG_OK) return rc;
    if (!h_out3) {
        gg_set_error("gg_overflow_info: null output");
        return GG_ERR_INVALID;
    }
    for (int i = 0; i < 3; ++i) h_out3[i] = ctx->last_overflow[i];
    return GG_OK;
}

int gg_last_batch_stats(gg_context *ctx, int n, int64_t *h_out) {
    int rc = check_ctx(ctx, false);
    if (rc != GG_OK) return rc;
    if (n > ctx->last_batch_n) n = ctx->last_batch_n;
    for (int i = 0; i < n; ++i) {
        int32_t c[8];
        GG_CUDA(cudaMemcpy(c, ctx->vset[ctx->cur].v[i].counters, sizeof(c), cudaMemcpyDeviceToHost));
        h_out[4 * i + 0] = c[0];
        h_out[4 * i + 1] = c[6] > c[1] ? c[6] : c[1];  // face records wanted, even beyond the capacity
        h_out[4 * i + 2] = c[2];
        h_out[4 * i + 3] = c[3];
    }
    return GG_OK;
}

int gg_set_mesh(gg_context *ctx, const float *d_verts, int64_t V, const int32_t *d_faces, int64_t F, void *stream) {
    int rc = check_ctx(ctx, false);
    if (rc != GG_OK) return rc;
    if (!d_verts || !d_faces || V < 1 || F < 1 || V >= (1LL << 31) || F >= (1LL << 31)) {
        gg_set_error("gg_set_mesh: need 1 <= V,F < 2^31 and non-null device pointers");
        return GG_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    GG_CUDA(cudaDeviceSynchronize());
    cudaFree(ctx->d_verts);
    cudaFree(ctx->d_faces);
    cudaFree(ctx->d_block_lo);
    cudaFree(ctx->d_block_hi);
    cudaFree(ctx->d_winner);
    cudaFree(ctx->d_wdense);
    // the scratch slots are laid out for THIS mesh (visible-block lists are sized by its block count): start over
    cudaFree(ctx->d_scratch);
    ctx->d_scratch = nullptr;
    ctx->scratch_bytes = 0;
    ctx->n_slots = 0;
    ctx->slot_tiles = 0;
    ctx->cap_recs = ctx->cap_bins = 0;
    for (int s = 0; s < GG_NSETS; ++s) ctx->ras_pending[s] = false;
    ctx->d_wdense = nullptr;
    ctx->wdense_cap = 0;
    ctx->d_verts = nullptr;
    ctx->d_faces = nullptr;
    ctx->d_block_lo = ctx->d_block_hi = nullptr;
    ctx->d_winner = nullptr;
    ctx->winner_cap = 0;
    ctx->F = 0;
    ctx->V = 0;
    ctx->n_blocks = (F + GG_BLOCK_FACES - 1) / GG_BLOCK_FACES;
    GG_CUDA(cudaMalloc(&ctx->d_verts, (size_t)V * sizeof(float4)));
    GG_CUDA(cudaMalloc(&ctx->d_faces, (size_t)F * sizeof(int4)));
    GG_CUDA(cudaMalloc(&ctx->d_block_lo, (size_t)ctx->n_blocks * 3 * sizeof(float)));
    GG_CUDA(cudaMalloc(&ctx->d_block_hi, (size_t)ctx->n_blocks * 3 * sizeof(float)));
    int *d_bad = nullptr;
    GG_CUDA(cudaMalloc(&d_bad, sizeof(int)));
    GG_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
    const int64_t n = V > F ? V : F;
    GG_LAUNCH(ctx, GG_ST_MESH, st,
              k_pack_mesh<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_verts, V, d_faces, F, ctx->d_verts, ctx->d_faces, d_bad));
    int bad = 0;
    GG_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    GG_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_bad);
    if (bad) {
        gg_set_error("gg_set_mesh: face index out of range");
        return GG_ERR_INVALID;
    }
    ctx->V = V;
    ctx->F = F;
    rc = sort_faces_spatially(ctx, st);
    if (rc != GG_OK) return rc;
    rc = gg_launch_mesh_blocks(ctx, st);
    if (rc != GG_OK) return rc;
    GG_CUDA(cudaStreamSynchronize(st));
    return GG_OK;
}

int gg_project(gg_context *ctx, const gg_camera *h_cams, int n, int32_t *d_X, int32_t *d_Y, float *d_invz,
               uint8_t *d_valid, void *stream) {
    int rc = check_ctx(ctx, true);
    if (rc != GG_OK) return rc;
    if (!h_cams || n < 1 || n > GG_MAX_VIEWS_PER_CALL || !d_X || !d_Y || !d_invz || !d_valid) {
        gg_set_error("gg_project: bad arguments");
        return GG_ERR_INVALID;
    }
    rc = gg_pipeline_drain(ctx, (cudaStream_t)stream);
    if (rc != GG_OK) return rc;
    return gg_launch_project(ctx, h_cams, n, d_X, d_Y, d_invz, d_valid, (cudaStream_t)stream);
}

int gg_rasterize(gg_context *ctx, const gg_camera *h_cams, int n, int32_t *d_pix2face, float *d_depth, void *stream) {
    int rc = check_ctx(ctx, true);
    if (rc != GG_OK) return rc;
    rc = check_cams(h_cams, n);
    if (rc != GG_OK) return rc;
    if (!d_pix2face) {
        gg_set_error("gg_rasterize: d_pix2face is null");
        return GG_ERR_INVALID;
    }
    rc = gg_pipeline_drain(ctx, (cudaStream_t)stream);
    if (rc != GG_OK) return rc;
    return gg_launch_rasterize(ctx, h_cams, n, d_pix2face, d_depth, 0, 0, (cudaStream_t)stream, (cudaStream_t)stream);
}

int gg_rasterize_render_flat(gg_context *ctx, const gg_camera *h_cams, int n, const double *d_face_tex, int D,
                             void *d_out, int out_dtype, int32_t *d_pix2face, void *stream) {
    int rc = check_ctx(ctx, true);
    if (rc != GG_OK) return rc;
    rc = check_cams(h_cams, n);
    if (rc != GG_OK) return rc;
    if (!d_face_tex || !d_out || D < 1) {
        gg_set_error("gg_rasterize_render_flat: bad arguments");
        return GG_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    rc = gg_pipeline_drain(ctx, st);
    if (rc != GG_OK) return rc;
    return gg_launch_rasterize(ctx, h_cams, n, d_pix2face, nullptr, 0, 0, st, st, nullptr, 0, 0, nullptr, nullptr, d_face_tex,
                               D, d_out, out_dtype);
}

int gg_aggregate(gg_context *ctx, const int32_t *d_pix2face, int H, int W, const void *d_pred, int pred_kind, int C,
                 int mode, int flags, double *d_sum, int32_t *d_count, void *stream) {
    int rc = check_ctx(ctx, true);
    if (rc != GG_OK) return rc;
    if (!d_pix2face || !d_pred || !d_sum || !d_count || H < 1 || W < 1 || C < 1) {
        gg_set_error("gg_aggregate: bad arguments");
        return GG_ERR_INVALID;
    }
    rc = gg_pipeline_drain(ctx, (cudaStream_t)stream);
    if (rc != GG_OK) return rc;
    return gg_launch_aggregate(ctx, d_pix2face, H, W, d_pred, pred_kind, C, mode, flags, d_sum, d_count,
                               (cudaStream_t)stream);
}

int gg_project_aggregate(gg_context *ctx, const gg_camera *h_cams, int n, const void *const *h_pred, int pred_kind,
                         int C, int mode, int flags, double *d_sum, int32_t *d_count, int32_t *d_pix2face,
                         void *stream) {
    int rc = check_ctx(ctx, true);
    if (rc != GG_OK) return rc;
    rc = check_cams(h_cams, n);
    if (rc != GG_OK) return rc;
    if (!h_pred || !d_sum || !d_count || C < 1) {
        gg_set_error("gg_project_aggregate: bad arguments");
        return GG_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int W = h_cams[0].W, H = h_cams[0].H;
    const int64_t P = (int64_t)W * H;
    if (P >= (1LL << 31)) {
        gg_set_error("gg_project_aggregate: raster larger than 2^31 pixels");
        return GG_ERR_INVALID;
    }
    for (int i = 0; i < n; ++i) {  // the kernels dereference these: refuse what the GPU cannot read
        cudaPointerAttributes attr;
        const bool ok = h_pred[i] && cudaPointerGetAttributes(&attr, h_pred[i]) == cudaSuccess &&
                        (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeHost ||
                         attr.type == cudaMemoryTypeManaged);
        (void)cudaGetLastError();
        if (!ok) {
            gg_set_error("gg_project_aggregate: prediction image is not in device or page-locked host memory "
                         "(for pageable arrays use gg_project_winners + gg_accumulate_rows)");
            return GG_ERR_INVALID;
        }
    }
    const bool fused = (mode == GG_MODE_LAST_PIXEL || mode == GG_MODE_VOTE) || (mode == GG_MODE_PIXEL_SUM && C <= 32);
    if (fused) {
        // Fused: the rasters never touch HBM unless the caller asked for them.  Last-pixel / vote: the rasterizer
        // leaves every face's last pixel in scratch and k_resolve_batch applies the views in order (bit-identical
        // to the reference's loop, meshes.py:2056-2062).  Pixel-sum: the rasterizer's dense epilogue streams the
        // scores.  Software pipeline: this batch is binned on stream A while the previous batch is still being
        // rasterized on stream B; the batches rotate through GG_NSETS sets of scratch slots.  The accumulators are
        // complete once gg_finalize / gg_sync (or any other entry point) has been called on the caller's stream.
        cudaStream_t sb = st, sr = st;
        if (ctx->pipeline) {
            sb = ctx->sA;
            sr = ctx->sB;
            ctx->cur = ctx->parity;
            ctx->parity = (ctx->parity + 1) % GG_NSETS;
            GG_CUDA(cudaEventRecord(ctx->ev_user, st));
            GG_CUDA(cudaStreamWaitEvent(sb, ctx->ev_user, 0));  // after whatever the caller enqueued before
            GG_CUDA(cudaStreamWaitEvent(sr, ctx->ev_user, 0));
            if (ctx->ras_pending[ctx->cur]) GG_CUDA(cudaStreamWaitEvent(sb, ctx->ev_ras[ctx->cur], 0));  // slot set free?
        } else {
            rc = gg_pipeline_drain(ctx, st);
            if (rc != GG_OK) return rc;
        }
        if (mode == GG_MODE_PIXEL_SUM) {
            rc = gg_launch_rasterize(ctx, h_cams, n, d_pix2face, nullptr, 0, 0, sb, sr, h_pred, pred_kind, C, d_sum, d_count);
        } else {
            rc = gg_launch_rasterize(ctx, h_cams, n, d_pix2face, nullptr, 1, flags & GG_FLAG_COMPAT_NEG, sb, sr);
            if (rc != GG_OK) return rc;
            if (ctx->pipeline) {  // the (latency-bound) resolve gets its own stream: the next rasterizer need not wait
                GG_CUDA(cudaEventRecord(ctx->ev_mid[ctx->cur], sr));
                sr = ctx->sC;
                GG_CUDA(cudaStreamWaitEvent(sr, ctx->ev_mid[ctx->cur], 0));
            }
            rc = gg_launch_resolve_batch(ctx, n, h_pred, pred_kind, C, mode, flags, d_sum, d_count, sr);
        }
        if (rc != GG_OK) return rc;
        if (ctx->pipeline) {
            GG_CUDA(cudaEventRecord(ctx->ev_ras[ctx->cur], sr));
            ctx->ras_pending[ctx->cur] = true;
        }
        return GG_OK;
    }
    rc = gg_pipeline_drain(ctx, st);
    if (rc != GG_OK) return rc;
    int32_t *raster = d_pix2face;
    if (!raster) {
        const int64_t need = P * n;
        if (ctx->raster_cap < need) {
            if (ctx->d_raster) {
                GG_CUDA(cudaDeviceSynchronize());
                GG_CUDA(cudaFree(ctx->d_raster));
                ctx->d_raster = nullptr;
            }
            GG_CUDA(cudaMalloc(&ctx->d_raster, (size_t)need * 4));
            ctx->raster_cap = need;
        }
        raster = ctx->d_raster;
    }
    rc = gg_launch_rasterize(ctx, h_cams, n, raster, nullptr, 0, 0, st, st);
    if (rc != GG_OK) return rc;
    for (int i = 0; i < n; ++i) {
        rc = gg_launch_aggregate(ctx, raster + P * i, H, W, h_pred[i], pred_kind, C, mode, flags, d_sum, d_count, st);
        if (rc != GG_OK) return rc;
    }
    return GG_OK;
}

int gg_project_winners(gg_context *ctx, const gg_camera *h_cams, int n, int flags, int32_t *d_pairs,
                       int64_t cap_pairs_per_view, int32_t *d_counts, void *stream) {
    int rc = check_ctx(ctx, true);
    if (rc != GG_OK) return rc;
    rc = check_cams(h_cams, n);
    if (rc != GG_OK) return rc;
    if (!d_pairs || !d_counts || cap_pairs_per_view < 1) {
        gg_set_error("gg_project_winners: bad arguments");
        return GG_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    rc = gg_pipeline_drain(ctx, st);
    if (rc != GG_OK) return rc;
    rc = gg_launch_rasterize(ctx, h_cams, n, nullptr, nullptr, 1, flags & GG_FLAG_COMPAT_NEG, st, st);
    if (rc != GG_OK) return rc;
    return gg_launch_compact_winners(ctx, n, flags, d_pairs, cap_pairs_per_view, d_counts, st);
}

int gg_accumulate_rows(gg_context *ctx, const int32_t *d_pairs, int64_t n_rows, const void *d_rows, int pred_kind, int C,
                       int mode, int flags, double *d_sum, int32_t *d_count, void *stream) {
    int rc = check_ctx(ctx, true);
    if (rc != GG_OK) return rc;
    if (n_rows < 0 || (n_rows > 0 && (!d_pairs || !d_rows)) || !d_sum || !d_count || C < 1 ||
        (mode != GG_MODE_LAST_PIXEL && mode != GG_MODE_VOTE)) {
        gg_set_error("gg_accumulate_rows: bad arguments (modes: GG_MODE_LAST_PIXEL, GG_MODE_VOTE)");
        return GG_ERR_INVALID;
    }
    rc = gg_pipeline_drain(ctx, (cudaStream_t)stream);
    if (rc != GG_OK) return rc;
    return gg_launch_accumulate_rows(ctx, d_pairs, n_rows, d_rows, pred_kind, C, mode, flags, d_sum, d_count,
                                     (cudaStream_t)stream);
}

int gg_finalize(gg_context *ctx, double *d_sum, const int32_t *d_count, int64_t F, int C, double *d_avg,
                double *d_argmax, void *stream) {
    int rc = check_ctx(ctx, false);
    if (rc != GG_OK) return rc;
    if (!d_sum || !d_count || F < 1 || C < 1) {
        gg_set_error("gg_finalize: bad arguments");
        return GG_ERR_INVALID;
    }
    rc = gg_pipeline_drain(ctx, (cudaStream_t)stream);
    if (rc != GG_OK) return rc;
    return gg_launch_finalize(ctx, d_sum, d_count, F, C, d_avg, d_argmax, (cudaStream_t)stream);
}

int gg_render_flat(gg_context *ctx, const int32_t *d_pix2face, int64_t n_pixels, const double *d_face_tex, int D,
                   void *d_out, int out_dtype, void *stream) {
    int rc = check_ctx(ctx, false);
    if (rc != GG_OK) return rc;
    if (!d_pix2face || !d_face_tex || !d_out || n_pixels < 1 || D < 1) {
        gg_set_error("gg_render_flat: bad arguments");
        return GG_ERR_INVALID;
    }
    rc = gg_pipeline_drain(ctx, (cudaStream_t)stream);
    if (rc != GG_OK) return rc;
    return gg_launch_render_flat(ctx, d_pix2face, n_pixels, d_face_tex, D, d_out, out_dtype, (cudaStream_t)stream);
}

}  // extern "C"
