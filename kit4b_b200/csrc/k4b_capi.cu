// k4b_capi.cu - extern "C" boundary of the engine (include/k4b_hamm.h): device management,
// packing, query sharding over the GPUs of one box, NCCL broadcast of the packed target set,
// gather of per-GPU minima.  Host orchestration only; the arithmetic is in k4b_kernels.cu.
// There is deliberately no CPU implementation behind any entry point.
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <stdlib.h>

#include <algorithm>
#include <chrono>
#include <mutex>
#include <cmath>
#include <vector>

#include "../../include/k4b_hamm.h"
#include "k4b_kernels.cuh"

using namespace k4b;

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
static int cuda_code(cudaError_t e) {
    return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? K4B_ERR_NODEVICE
           : (e == cudaErrorMemoryAllocation)                           ? K4B_ERR_MEM
                                                                        : K4B_ERR_CUDA;
}
#define CU(call)                                                                               \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(cuda_code(e_), "%s:%d %s: %s", __FILE__, __LINE__, #call,              \
                        cudaGetErrorString(e_));                                               \
    } while (0)
#define RC(call)              \
    do {                      \
        int rc_ = (call);     \
        if (rc_) return rc_;  \
    } while (0)

// ------------------------------------------------------------------------------------------
// NCCL through dlopen (only needed when ONE process drives several GPUs; under torchrun the
// broadcast is issued by torch.distributed on the image pointer instead)
// ------------------------------------------------------------------------------------------
typedef struct ncclComm *ncclComm_t;
struct NcclApi {
    void *h = nullptr;
    int (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int /*ncclDataType_t*/, int /*root*/,
                     ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int /*ncclDataType_t*/, int /*ncclRedOp_t*/,
                     ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
constexpr int kNcclUint8 = 1;   // ncclUint8 in nccl.h
constexpr int kNcclUint32 = 3;  // ncclUint32
constexpr int kNcclMin = 3;     // ncclMin
constexpr int kNcclMax = 2;     // ncclMax

struct Engine {
    std::vector<int> devs;
    std::vector<cudaStream_t> streams;
    std::vector<ncclComm_t> comms;
    // pinned host staging for results, one grow-only buffer per device slot: cudaMallocHost costs
    // milliseconds (hundreds for a gigabyte), the host-buffer entry points would pay it per call
    std::vector<void *> pinned;
    std::vector<size_t> pinned_bytes;
    NcclApi nccl;
    bool inited = false;
};
static Engine g_eng;
static std::mutex g_mu;

// Stream-ordered device scratch from the retained default pool (k4b_gpu_init sets the release
// threshold to "never"): after the first call of a given size an allocation is a pointer bump, and
// the destructor returns the block on every path, early error returns included.
struct DevScratch {
    void *p = nullptr;
    cudaStream_t st = nullptr;
    DevScratch() = default;
    DevScratch(const DevScratch &) = delete;
    DevScratch &operator=(const DevScratch &) = delete;
    ~DevScratch() { release(); }
    cudaError_t alloc(size_t bytes, cudaStream_t stream) {
        release();
        st = stream;
        return cudaMallocAsync(&p, bytes ? bytes : 16, stream);
    }
    void release() {
        if (p) cudaFreeAsync(p, st);
        p = nullptr;
    }
    template <typename T>
    T *as() const { return static_cast<T *>(p); }
};

static void *pinned_result(int slot, size_t bytes) {
    if ((size_t)slot >= g_eng.pinned.size()) return nullptr;
    if (g_eng.pinned_bytes[slot] < bytes) {
        if (g_eng.pinned[slot]) cudaFreeHost(g_eng.pinned[slot]);
        g_eng.pinned[slot] = nullptr;
        g_eng.pinned_bytes[slot] = 0;
        const size_t want = bytes + bytes / 8 + 4096;
        if (cudaMallocHost(&g_eng.pinned[slot], want) != cudaSuccess) {
            cudaGetLastError();
            g_eng.pinned[slot] = nullptr;
            return nullptr;
        }
        g_eng.pinned_bytes[slot] = want;
    }
    return g_eng.pinned[slot];
}

static int load_nccl(NcclApi &n) {
    if (n.h) return 0;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        n.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (n.h) break;
    }
    if (!n.h) return fail(K4B_ERR_NCCL, "dlopen libnccl.so.2 failed: %s", dlerror());
#define SYM(field, name)                     \
    *(void **)(&n.field) = dlsym(n.h, name); \
    if (!n.field) return fail(K4B_ERR_NCCL, "libnccl lacks symbol %s", name)
    SYM(CommInitAll, "ncclCommInitAll");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(Broadcast, "ncclBroadcast");
    SYM(AllReduce, "ncclAllReduce");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    return 0;
}

extern "C" int k4b_gpu_init(int n_gpus, const int *device_ids) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_eng.inited) return K4B_OK;
    int avail = 0;
    cudaError_t e = cudaGetDeviceCount(&avail);
    if (e != cudaSuccess || avail <= 0)
        return fail(K4B_ERR_NODEVICE, "no CUDA device: %s (this engine has no CPU fallback)",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (n_gpus <= 0) n_gpus = avail;
    if (n_gpus > avail) return fail(K4B_ERR_PARAMS, "%d GPUs requested, %d visible", n_gpus, avail);
    g_eng.devs.clear();
    for (int i = 0; i < n_gpus; ++i) {
        const int d = device_ids ? device_ids[i] : i;
        if (d < 0 || d >= avail) return fail(K4B_ERR_PARAMS, "bad device id %d", d);
        g_eng.devs.push_back(d);
    }
    g_eng.streams.assign(n_gpus, nullptr);
    g_eng.pinned.assign(n_gpus, nullptr);
    g_eng.pinned_bytes.assign(n_gpus, 0);
    for (int i = 0; i < n_gpus; ++i) {
        CU(cudaSetDevice(g_eng.devs[i]));
        CU(cudaStreamCreateWithFlags(&g_eng.streams[i], cudaStreamNonBlocking));
        {   // keep stream-ordered allocations cached in the pool: the default (release threshold 0)
            // hands multi-GB scratch (seed index) back to the driver at the next synchronisation,
            // which stalled the host for seconds (profiles/r01_seed_engine.log)
            cudaMemPool_t pool = nullptr;
            if (cudaDeviceGetDefaultMemPool(&pool, g_eng.devs[i]) == cudaSuccess && pool) {
                unsigned long long keep = ~0ull;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
        }
    }
    if (n_gpus > 1) {
        RC(load_nccl(g_eng.nccl));
        g_eng.comms.assign(n_gpus, nullptr);
        int nr = g_eng.nccl.CommInitAll(g_eng.comms.data(), n_gpus, g_eng.devs.data());
        if (nr) return fail(K4B_ERR_NCCL, "ncclCommInitAll: %s", g_eng.nccl.GetErrorString(nr));
    }
    CU(cudaSetDevice(g_eng.devs[0]));
    g_eng.inited = true;
    return K4B_OK;
}

extern "C" void k4b_gpu_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_eng.inited) return;
    for (size_t i = 0; i < g_eng.comms.size(); ++i)
        if (g_eng.comms[i]) g_eng.nccl.CommDestroy(g_eng.comms[i]);
    g_eng.comms.clear();
    for (size_t i = 0; i < g_eng.devs.size(); ++i) {
        cudaSetDevice(g_eng.devs[i]);
        if (g_eng.streams[i]) {
            cudaStreamSynchronize(g_eng.streams[i]);
            cudaStreamDestroy(g_eng.streams[i]);
        }
        if (i < g_eng.pinned.size() && g_eng.pinned[i]) cudaFreeHost(g_eng.pinned[i]);
        cudaMemPool_t pool = nullptr;  // hand the retained scratch back to the driver
        if (cudaDeviceGetDefaultMemPool(&pool, g_eng.devs[i]) == cudaSuccess && pool) cudaMemPoolTrimTo(pool, 0);
    }
    g_eng.pinned.clear();
    g_eng.pinned_bytes.clear();
    g_eng.streams.clear();
    g_eng.devs.clear();
    g_eng.inited = false;
}

extern "C" int k4b_gpu_count(void) { return g_eng.inited ? (int)g_eng.devs.size() : 0; }
extern "C" const char *k4b_last_error(void) { return g_err; }

static int ensure_init() {
    if (g_eng.inited) return K4B_OK;
    return k4b_gpu_init(1, nullptr);
}

// ------------------------------------------------------------------------------------------
// packed images
// ------------------------------------------------------------------------------------------
struct k4b_packed {
    uint32_t *d_image = nullptr;
    bool owns = false;
    // pooled: image and reverse-complement planes come from the stream-ordered pool and go back to it
    // on pool_stream (the host-buffer entry points); otherwise cudaMalloc / cudaFree, which is safe
    // whatever stream the caller used the handle on
    bool pooled = false;
    cudaStream_t pool_stream = nullptr;
    int device = 0;
    uint32_t len = 0, K = 0, nw = 0, stride = 0;
    int has_non_acgt = 0;
    uint64_t num_kmers = 0;
    uint32_t *d_rc_planes = nullptr;  // lazily built: reverse-complemented planes (same geometry)
    ImageView view() const {
        return ImageView{d_image + kFrontPadWords, stride, stride - (uint32_t)kFrontPadWords, len};
    }
    ImageView rc_view() const {
        return ImageView{d_rc_planes + kFrontPadWords, stride, stride - (uint32_t)kFrontPadWords, len};
    }
};

// words per array: front pad + sequence rounded up to whole tiles + back pad
static uint32_t array_stride(uint32_t len) {
    const uint32_t nw = (len + 31) / 32;
    return kFrontPadWords + (nw + kTileGroups - 1) / kTileGroups * kTileGroups + kBackPadWords;
}
extern "C" size_t k4b_packed_image_bytes(uint32_t concat_len) {
    return (size_t)array_stride(concat_len) * kImageArrays * sizeof(uint32_t);
}
extern "C" void *k4b_packed_image_ptr(k4b_packed *p) { return p ? p->d_image : nullptr; }
extern "C" size_t k4b_packed_image_size(k4b_packed *p) { return p ? (size_t)p->stride * kImageArrays * 4 : 0; }
extern "C" int k4b_packed_has_non_acgt(k4b_packed *p) { return p ? p->has_non_acgt : 0; }
extern "C" uint64_t k4b_packed_num_kmers(k4b_packed *p) { return p ? p->num_kmers : 0; }
extern "C" void k4b_packed_free(k4b_packed *p) {
    if (!p) return;
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(p->device);
    if (p->pooled) {
        if (p->owns && p->d_image) cudaFreeAsync(p->d_image, p->pool_stream);
        if (p->d_rc_planes) cudaFreeAsync(p->d_rc_planes, p->pool_stream);
    } else {
        if (p->owns && p->d_image) cudaFree(p->d_image);
        if (p->d_rc_planes) cudaFree(p->d_rc_planes);  // valid for stream-ordered allocations too
    }
    cudaSetDevice(cur);
    delete p;
}

static int check_k(uint32_t K, uint32_t lo, uint32_t hi) {
    if (K < lo || K > hi) return fail(K4B_ERR_PARAMS, "K=%u outside %u..%u", K, lo, hi);
    return 0;
}
static int check_len(uint64_t len) {
    if (!len) return fail(K4B_ERR_PARAMS, "empty sequence");
    if (len > 0xffffffffull - 64ull * kTileGroups)
        return fail(K4B_ERR_PARAMS, "sequence of %llu bases exceeds the 32-bit position space",
                    (unsigned long long)len);
    return 0;
}

// pack + valid-start kernels into `image` (caller- or library-owned)
static int pack_into(const void *d_concat, uint32_t concat_len, uint32_t K, uint32_t *image,
                     bool owns, cudaStream_t st, k4b_packed **out) {
    k4b_packed *p = new k4b_packed;
    cudaGetDevice(&p->device);
    p->len = concat_len;
    p->K = K;
    p->nw = (concat_len + 31) / 32;
    p->stride = array_stride(concat_len);
    p->owns = owns;
    p->d_image = image;
    struct Scratch {
        uint32_t flags;
        uint32_t pad;
        unsigned long long count;
    };
    DevScratch scratch;
    Scratch h_s = {0, 0, 0};
    cudaError_t e = scratch.alloc(sizeof(Scratch), st);
    Scratch *d_s = scratch.as<Scratch>();
    if (e == cudaSuccess) e = cudaMemsetAsync(d_s, 0, sizeof(Scratch), st);
    if (e == cudaSuccess)
        e = launch_pack((const uint8_t *)d_concat, p->view(), &d_s->flags, st);
    if (e == cudaSuccess) e = launch_valid(p->view(), K, &d_s->count, st);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(&h_s, d_s, sizeof(Scratch), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    scratch.release();
    if (e != cudaSuccess) {
        k4b_packed_free(p);
        return fail(cuda_code(e), "pack (%u bases): %s", concat_len, cudaGetErrorString(e));
    }
    p->has_non_acgt = (int)(h_s.flags & 1u);
    p->num_kmers = h_s.count;
    *out = p;
    return K4B_OK;
}

static int pack_args_ok(const void *d_concat, uint32_t concat_len, uint32_t K, k4b_packed **out) {
    if (!out) return fail(K4B_ERR_PARAMS, "out is NULL");
    *out = nullptr;
    RC(ensure_init());
    RC(check_k(K, K4B_MIN_K, K4B_MAX_K));
    if (!d_concat) return fail(K4B_ERR_PARAMS, "NULL sequence");
    RC(check_len(concat_len));
    if (((uintptr_t)d_concat & 15) != 0)
        return fail(K4B_ERR_PARAMS, "device concat must be 16-byte aligned");
    return 0;
}

extern "C" int k4b_pack_device(const void *d_concat, uint32_t concat_len, uint32_t K, void *stream,
                               k4b_packed **out) {
    RC(pack_args_ok(d_concat, concat_len, K, out));
    uint32_t *image = nullptr;
    cudaError_t e = cudaMalloc(&image, k4b_packed_image_bytes(concat_len));
    if (e != cudaSuccess)
        return fail(cuda_code(e), "cudaMalloc packed image: %s", cudaGetErrorString(e));
    return pack_into(d_concat, concat_len, K, image, true, (cudaStream_t)stream, out);
}

extern "C" int k4b_pack_device_into(const void *d_concat, uint32_t concat_len, uint32_t K,
                                    void *d_image, size_t image_bytes, void *stream,
                                    k4b_packed **out) {
    RC(pack_args_ok(d_concat, concat_len, K, out));
    if (!d_image || image_bytes != k4b_packed_image_bytes(concat_len) || ((uintptr_t)d_image & 15))
        return fail(K4B_ERR_PARAMS, "image buffer must be 16-byte aligned and %zu bytes",
                    k4b_packed_image_bytes(concat_len));
    return pack_into(d_concat, concat_len, K, (uint32_t *)d_image, false, (cudaStream_t)stream, out);
}

// host concat -> packed image on engine device slot `slot`, everything from the stream-ordered pool
// on that device's engine stream (the host-buffer entry points).  A pinned source is copied at full
// PCIe rate; a pageable one is staged by the driver.
static int pack_host_pooled(const uint8_t *concat, uint32_t concat_len, uint32_t K, int slot, k4b_packed **out) {
    *out = nullptr;
    RC(check_k(K, K4B_MIN_K, K4B_MAX_K));
    RC(check_len(concat_len));
    CU(cudaSetDevice(g_eng.devs[slot]));
    cudaStream_t st = g_eng.streams[slot];
    DevScratch d_c;
    CU(d_c.alloc(((size_t)concat_len + 15) / 16 * 16, st));
    CU(cudaMemcpyAsync(d_c.p, concat, concat_len, cudaMemcpyHostToDevice, st));
    uint32_t *image = nullptr;
    CU(cudaMallocAsync(&image, k4b_packed_image_bytes(concat_len), st));
    RC(pack_into(d_c.p, concat_len, K, image, true, st, out));  // synchronises st; owns the image even on failure
    (*out)->pooled = true;
    (*out)->pool_stream = st;
    return K4B_OK;
}

extern "C" int k4b_pack_host(const uint8_t *concat, uint32_t concat_len, uint32_t K,
                             k4b_packed **out) {
    if (!out) return fail(K4B_ERR_PARAMS, "out is NULL");
    *out = nullptr;
    RC(ensure_init());
    if (!concat) return fail(K4B_ERR_PARAMS, "NULL sequence");
    RC(check_len(concat_len));
    uint8_t *d_c = nullptr;
    CU(cudaMalloc(&d_c, ((size_t)concat_len + 15) / 16 * 16));
    cudaError_t e = cudaMemcpy(d_c, concat, concat_len, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(d_c);
        return fail(K4B_ERR_CUDA, "H2D concat: %s", cudaGetErrorString(e));
    }
    int rc = k4b_pack_device(d_c, concat_len, K, nullptr, out);
    cudaFree(d_c);
    return rc;
}

extern "C" int k4b_packed_from_image(void *d_image, size_t image_bytes, uint32_t concat_len,
                                     uint32_t K, int has_non_acgt, k4b_packed **out) {
    if (!out) return fail(K4B_ERR_PARAMS, "out is NULL");
    *out = nullptr;
    if (!d_image || image_bytes != k4b_packed_image_bytes(concat_len))
        return fail(K4B_ERR_PARAMS, "image is %zu bytes, expected %zu", image_bytes,
                    k4b_packed_image_bytes(concat_len));
    RC(ensure_init());
    RC(check_k(K, K4B_MIN_K, K4B_MAX_K));
    k4b_packed *p = new k4b_packed;
    cudaGetDevice(&p->device);
    p->d_image = (uint32_t *)d_image;
    p->owns = false;
    p->len = concat_len;
    p->K = K;
    p->nw = (concat_len + 31) / 32;
    p->stride = array_stride(concat_len);
    p->has_non_acgt = has_non_acgt;
    p->num_kmers = 0;
    *out = p;
    return K4B_OK;
}

// ------------------------------------------------------------------------------------------
// all-pairs on device-resident data
// ------------------------------------------------------------------------------------------
// Kernel timing: CUDA event pairs recorded on the launching stream around the engine launches.
// A run may consist of several calls (band slabs with collectives in between): every call adds a
// pair, k4b_last_kernel_ms() sums the pairs recorded since the run began.
struct TimingPool {
    std::vector<cudaEvent_t> ev;  // 2 per pair
    size_t used = 0;              // pairs in use
    int dev = -1;
    int begin(int device, bool reset, cudaStream_t st) {
        if (dev != device) {
            for (cudaEvent_t e : ev) cudaEventDestroy(e);
            ev.clear();
            used = 0;
            dev = device;
        }
        if (reset) used = 0;
        if (ev.size() < 2 * (used + 1)) {
            cudaEvent_t a, b;
            CU(cudaEventCreate(&a));
            CU(cudaEventCreate(&b));
            ev.push_back(a);
            ev.push_back(b);
        }
        ++used;
        CU(cudaEventRecord(ev[2 * used - 2], st));
        return 0;
    }
    cudaError_t end(cudaStream_t st) { return used ? cudaEventRecord(ev[2 * used - 1], st) : cudaSuccess; }
    cudaEvent_t last() const { return used ? ev[2 * used - 1] : nullptr; }
};
static thread_local TimingPool g_tp;

extern "C" float k4b_last_kernel_ms(void) {
    if (!g_tp.used) return -1.f;
    float total = 0.f;
    for (size_t i = 0; i < g_tp.used; ++i) {
        float ms = 0.f;
        if (cudaEventSynchronize(g_tp.ev[2 * i + 1]) != cudaSuccess) return -1.f;
        if (cudaEventElapsedTime(&ms, g_tp.ev[2 * i], g_tp.ev[2 * i + 1]) != cudaSuccess) return -1.f;
        total += ms;
    }
    return total;
}

struct SweepRange {  // see AllPairsParams
    bool ranged = false;
    long long w_lo = 0, w_hi = 0, c1_lo = 0, c1_hi = -1, c2_lo = 0, c2_hi = -1;
};
static int allpairs_impl(k4b_packed *queries, k4b_packed *targets, int both_strands,
                         int self_exclude, uint32_t q_begin, uint32_t q_end, uint32_t clamp,
                         const SweepRange &sweep, uint16_t *d_out_min, void *stream, int *launches);

// -z of the call in progress on this thread (k4b_hamm_targeted_z): mode 1 intra / 2 inter and
// the flat start positions of the entries of the indexed assembly
struct ZFilter {
    int mode = 0;
    std::vector<uint32_t> starts;
};
static thread_local ZFilter g_zf;

extern "C" int k4b_allpairs_min_device(k4b_packed *queries, k4b_packed *targets, int both_strands,
                                       int self_exclude, uint32_t q_begin, uint32_t q_end,
                                       uint32_t clamp, uint16_t *d_out_min, void *stream,
                                       int *launches) {
    return allpairs_impl(queries, targets, both_strands, self_exclude, q_begin, q_end, clamp,
                         SweepRange(), d_out_min, stream, launches);
}

static int allpairs_impl(k4b_packed *queries, k4b_packed *targets, int both_strands,
                         int self_exclude, uint32_t q_begin, uint32_t q_end, uint32_t clamp,
                         const SweepRange &sweep, uint16_t *d_out_min, void *stream, int *launches) {
    // clamp > 0 selects the targeted (-m0) rules: results capped at the "not found" value, probe
    // symbols >= N are wildcards, more than 4 of them report 0
    const bool targeted_rules = clamp > 0;
    if (launches) *launches = 0;
    if (!queries || !targets) return fail(K4B_ERR_PARAMS, "NULL packed handle");
    if (queries->K != targets->K) return fail(K4B_ERR_PARAMS, "query/target K differ");
    if (queries->device != targets->device)
        return fail(K4B_ERR_PARAMS, "query and target images live on different devices");
    if (q_end > queries->len) q_end = queries->len;
    if (q_begin >= q_end) return K4B_OK;
    if (!d_out_min) return fail(K4B_ERR_PARAMS, "d_out_min is NULL");
    CU(cudaSetDevice(queries->device));
    const uint32_t K = queries->K;
    const uint32_t nq = q_end - q_begin;
    cudaStream_t st = (cudaStream_t)stream;
    const bool three = queries->has_non_acgt || targets->has_non_acgt;
    const bool crick = both_strands != 0;
    const uint32_t W = (K + 31) / 32;
    const bool generic = W > (uint32_t)kMaxRegW;

    if (generic && crick && !queries->d_rc_planes) {
        CU(cudaMallocAsync(&queries->d_rc_planes, (size_t)queries->stride * kImageArrays * 4, st));
        CU(launch_revcomp_planes(queries->view(), queries->rc_view(), st));
    }

    DevScratch min_buf, ent_buf;  // returned to the pool on every path
    CU(min_buf.alloc((size_t)nq * 4, st));
    uint32_t *d_min32 = min_buf.as<uint32_t>();
    CU(launch_fill_u32(d_min32, nq, K + 1, st));

    AllPairsParams prm;
    prm.q = queries->view();
    prm.t = targets->view();
    prm.K = K;
    prm.q_begin = q_begin;
    prm.q_end = q_end;
    prm.tiles_total = (targets->nw + kTileGroups - 1) / kTileGroups;
    prm.groups_limit = 0;
    prm.out = d_min32;
    prm.self_exclude = self_exclude ? 1 : 0;
    prm.wildcard = targeted_rules ? 1 : 0;
    prm.zfilt = 0;
    prm.ent_starts = nullptr;
    prm.n_ent = 0;
    if (g_zf.mode && targeted_rules && self_exclude && queries == targets && !g_zf.starts.empty()) {
        CU(ent_buf.alloc(g_zf.starts.size() * 4, st));
        CU(cudaMemcpyAsync(ent_buf.p, g_zf.starts.data(), g_zf.starts.size() * 4, cudaMemcpyHostToDevice, st));
        prm.zfilt = g_zf.mode;
        prm.ent_starts = ent_buf.as<uint32_t>();
        prm.n_ent = (uint32_t)g_zf.starts.size();
    }
    prm.ranged = sweep.ranged ? 1 : 0;
    prm.w_lo = sweep.w_lo; prm.w_hi = sweep.w_hi;
    prm.c1_lo = sweep.c1_lo; prm.c1_hi = sweep.c1_hi;
    prm.c2_lo = sweep.c2_lo; prm.c2_hi = sweep.c2_hi;
    // enough CTAs for ~16 balanced waves on 148 SMs x 2 resident CTAs; never below one tile
    const uint32_t qpt = generic ? 1u : (uint32_t)queries_per_thread(W, three);
    const uint32_t qblocks = (nq + kThreads * qpt - 1) / (kThreads * qpt);
    uint32_t want_chunks = (4736 + qblocks - 1) / qblocks;
    want_chunks = std::max(1u, std::min(want_chunks, prm.tiles_total));
    prm.tiles_per_chunk = (prm.tiles_total + want_chunks - 1) / want_chunks;

    RC(g_tp.begin(queries->device, true, st));
    cudaError_t e = cudaSuccess;
    {
        if (!generic)
            e = launch_allpairs(prm, three, crick, st, nullptr);
        else
            e = launch_allpairs_generic(prm, three, crick,
                                        crick ? queries->rc_view() : queries->view(), st, nullptr);
    }
    if (e == cudaSuccess) e = g_tp.end(st);
    if (e == cudaSuccess)
        e = launch_finalize(d_min32, queries->view(), q_begin, nq, K, clamp, targeted_rules ? 4 : -1,
                            d_out_min, st);
    if (e != cudaSuccess) return fail(cuda_code(e), "allpairs launch: %s", cudaGetErrorString(e));
    if (launches) *launches = 3;  // fill + allpairs + finalize
    return K4B_OK;
}

// ------------------------------------------------------------------------------------------
// diagonal-band engine on device-resident data
// ------------------------------------------------------------------------------------------
// 0 auto, 1 POPC all-pairs, 2 diagonal bands, 3 seed-and-verify where it applies (targeted mode
// with pure-ACGT probes) else as auto; -1 = read K4B_ENGINE
static int g_engine = -1;
static int engine_setting() {
    if (g_engine < 0) {
        const char *e = getenv("K4B_ENGINE");
        g_engine = !e ? 0 : (!strcmp(e, "popc") ? 1 : (!strcmp(e, "diag") ? 2 : (!strcmp(e, "seed") ? 3 : 0)));
    }
    return g_engine;
}
extern "C" int k4b_set_engine(int engine) {
    if (engine < 0 || engine > 3)
        return fail(K4B_ERR_PARAMS, "engine must be 0 (auto), 1 (popc), 2 (diag) or 3 (seed)");
    g_engine = engine;
    return K4B_OK;
}
extern "C" int k4b_get_engine(void) { return engine_setting(); }

extern "C" int k4b_best_init_device(uint32_t *d_best, uint32_t n, uint32_t K, void *stream) {
    if (!d_best) return fail(K4B_ERR_PARAMS, "d_best is NULL");
    CU(launch_fill_u32(d_best, n, K + 1, (cudaStream_t)stream));
    return K4B_OK;
}

extern "C" int k4b_best_finalize_device(k4b_packed *g, const uint32_t *d_best, uint16_t *d_out_min,
                                        void *stream) {
    if (!g || !d_best || !d_out_min) return fail(K4B_ERR_PARAMS, "NULL argument");
    CU(cudaSetDevice(g->device));
    CU(launch_finalize(d_best, g->view(), 0, g->len, g->K, 0, -1, d_out_min, (cudaStream_t)stream));
    return K4B_OK;
}

static int diag_prepare(k4b_packed *g, bool crick, cudaStream_t st, int *nl) {
    CU(cudaSetDevice(g->device));
    if (crick && !g->d_rc_planes) {
        CU(cudaMallocAsync(&g->d_rc_planes, (size_t)g->stride * kImageArrays * 4, st));
        CU(launch_revcomp_planes(g->view(), g->rc_view(), st));
        ++*nl;
    }
    return 0;
}

// test hook: the image (6 arrays of stride words) and, when rc != 0, the lazily built
// reverse-complemented image of a handle, copied to the host
extern "C" int k4b_debug_copy_image(k4b_packed *g, int rc, uint32_t *host_out, size_t words) {
    if (!g || !host_out) return fail(K4B_ERR_PARAMS, "NULL argument");
    if (words != (size_t)g->stride * kImageArrays) return fail(K4B_ERR_PARAMS, "expected %zu words", (size_t)g->stride * kImageArrays);
    int nl = 0;
    if (rc) RC(diag_prepare(g, true, nullptr, &nl));
    CU(cudaSetDevice(g->device));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(host_out, rc ? g->d_rc_planes : g->d_image, words * 4, cudaMemcpyDeviceToHost));
    return K4B_OK;
}

// Bootstrap of the band engine: the K-mers starting in [q_begin, q_end) against a small sample
// of targets with the POPC engine, so that the thresholds of the band kernel start near the
// final minima.  Shards of queries are independent (combine d_best by element-wise minimum).
extern "C" int k4b_diag_bootstrap_device(k4b_packed *g, int both_strands, uint32_t q_begin,
                                         uint32_t q_end, uint32_t *d_best, void *stream) {
    if (!g || !d_best) return fail(K4B_ERR_PARAMS, "NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t K = g->K, len = g->len;
    if (q_end > len) q_end = len;
    if (len < K || q_begin >= q_end) return K4B_OK;
    const bool three = g->has_non_acgt != 0, crick = both_strands != 0;
    int nl = 0;
    RC(diag_prepare(g, crick, st, &nl));
    const uint32_t W = (K + 31) / 32;
    AllPairsParams bp;
    bp.q = g->view();
    bp.t = g->view();
    bp.K = K;
    bp.q_begin = q_begin;
    bp.q_end = q_end;
    const uint32_t tiles_all = (g->nw + kTileGroups - 1) / kTileGroups;
    const char *bt = getenv("K4B_BOOT_TILES");
    const uint32_t boot_tiles = bt ? (uint32_t)atoi(bt) : 1u;  // 8192 candidate starts; with the small first slabs
                                                               // one tile is enough (profiles/r01_slab_schedule.log)
    bp.tiles_total = std::min(tiles_all, std::max(1u, boot_tiles));
    // K > 128 runs the (slower) generic POPC kernel, and long K-mers have relatively narrow distance
    // distributions: a quarter tile (2048 candidate starts) bounds the minima almost as well
    bp.groups_limit = (W > (uint32_t)kMaxRegW && !bt) ? 64u : 0u;
    bp.out = d_best + q_begin;  // the kernel indexes out[] relative to q_begin
    bp.self_exclude = 1;
    bp.wildcard = 0;
    bp.ranged = 0;
    bp.zfilt = 0;
    bp.ent_starts = nullptr;
    bp.n_ent = 0;
    bp.w_lo = bp.w_hi = bp.c1_lo = bp.c1_hi = bp.c2_lo = bp.c2_hi = 0;
    const bool generic = W > (uint32_t)kMaxRegW;
    const uint32_t qpt = generic ? 1u : (uint32_t)queries_per_thread(W, three);
    const uint32_t nq = q_end - q_begin;
    const uint32_t qblocks = (nq + kThreads * qpt - 1) / (kThreads * qpt);
    const uint32_t want = std::max(1u, std::min((4736 + qblocks - 1) / qblocks, bp.tiles_total));
    bp.tiles_per_chunk = (bp.tiles_total + want - 1) / want;
    const cudaError_t e = generic ? launch_allpairs_generic(bp, three, crick, crick ? g->rc_view() : g->view(), st, nullptr)
                                  : launch_allpairs(bp, three, crick, st, nullptr);
    if (e != cudaSuccess) return fail(cuda_code(e), "bootstrap launch: %s", cudaGetErrorString(e));
    return K4B_OK;
}

// bookkeeping of the most recent band run of this thread (which counter width each slab used)
constexpr uint32_t kMaxSlabs = 64;  // slots per bookkeeping array (d_tmax, d_low on the device; the halves of g_h_tmax)
static thread_local uint32_t *g_h_tmax = nullptr;  // pinned, 2 * kMaxSlabs entries
static thread_local uint32_t g_info_slabs = 0, g_info_np_full = 0, g_info_np_small = 0, g_info_limit = 0,
                             g_info_low_max = 0;

extern "C" int k4b_last_diag_info(uint32_t *np_full, uint32_t *np_small, uint32_t *slabs,
                                  uint32_t *narrow_slabs) {
    if (!g_tp.used || !g_h_tmax) return fail(K4B_ERR_PARAMS, "no band run recorded on this thread");
    CU(cudaEventSynchronize(g_tp.last()));
    CU(cudaDeviceSynchronize());
    uint32_t narrow = 0;
    for (uint32_t i = 0; i < g_info_slabs; ++i)
        if (g_info_np_small && g_h_tmax[i] <= g_info_limit && g_h_tmax[kMaxSlabs + i] <= g_info_low_max) ++narrow;
    if (np_full) *np_full = g_info_np_full;
    if (np_small) *np_small = g_info_np_small;
    if (slabs) *slabs = g_info_slabs;
    if (narrow_slabs) *narrow_slabs = narrow;
    return K4B_OK;
}

// Slab schedule of the band engine.  A part owns the CTA groups g = part + nparts * q of each
// strand; slab j takes the q with (q mod R) in [bounds[j], bounds[j+1]).  The thresholds (block
// maxima of the running minima) are refreshed before every slab, so the first slabs are SMALL
// (1/R, 1/R, 2/R, 4/R ... of the work, then R/8 each): thresholds get tight after a few per cent
// of the pairs instead of after the first 1/16 (profiles/r01_slab_schedule.log).
namespace {
struct SlabPlan {
    uint32_t R = 1;
    std::vector<uint32_t> bounds;  // n_slabs + 1 entries, 0 .. R
    uint32_t n_slabs() const { return (uint32_t)bounds.size() - 1; }
};
SlabPlan make_slab_plan(uint64_t groups_per_part) {
    SlabPlan p;
    const char *sl = getenv("K4B_DIAG_SLABS");  // experiments: n equal slabs
    if (sl && atoi(sl) > 0) {
        p.R = (uint32_t)std::min((int)kMaxSlabs, atoi(sl));
        for (uint32_t i = 0; i <= p.R; ++i) p.bounds.push_back(i);
        return p;
    }
    const char *pr = getenv("K4B_DIAG_PERIOD"), *cp = getenv("K4B_DIAG_CAP");
    // at most kMaxSlabs slabs exist (per-slab bookkeeping slots of d_bm and g_h_tmax): the period,
    // which bounds the slab count, is clamped to it
    uint32_t rmax = pr ? (uint32_t)std::min(std::max(atoi(pr), 1), (int)kMaxSlabs) : kMaxSlabs;
    while (p.R < rmax && (uint64_t)p.R * 8 <= groups_per_part) p.R *= 2;
    const uint32_t cap = cp ? std::max(1, atoi(cp)) : std::max(1u, p.R / 8);
    p.bounds.push_back(0);
    for (uint32_t pos = 0; pos < p.R;) {
        pos += std::max(1u, std::min(std::min(cap, pos), p.R - pos));
        p.bounds.push_back(pos);
    }
    return p;
}
struct BandGeometry {
    uint32_t M = 0;
    uint64_t gw = 0, gc = 0;  // CTA groups of the Watson (s = 1..M) and Crick (s = -M..M) diagonals
    SlabPlan plan;
};
BandGeometry band_geometry(const k4b_packed *g, bool crick, uint32_t nparts) {
    BandGeometry bg;
    bg.M = g->len - g->K;
    bg.gw = ((uint64_t)bg.M + kDiagGroupDiagonals - 1) / kDiagGroupDiagonals;
    bg.gc = crick ? (2ull * bg.M + 1 + kDiagGroupDiagonals - 1) / kDiagGroupDiagonals : 0;
    bg.plan = make_slab_plan((bg.gw + bg.gc + nparts - 1) / nparts);
    return bg;
}
}  // namespace

extern "C" int k4b_diag_slab_count(k4b_packed *g, int both_strands, uint32_t nparts, uint32_t *n_slabs) {
    if (!g || !n_slabs || !nparts) return fail(K4B_ERR_PARAMS, "NULL argument");
    *n_slabs = g->len < g->K ? 0u : band_geometry(g, both_strands != 0, nparts).plan.n_slabs();
    return K4B_OK;
}

// A second stream per device for the Crick launches of a slab: Watson and Crick kernels of one slab
// are independent, so the CTAs of one fill the SMs that the draining tail of the other leaves idle
// (one partial wave per slab instead of two - it matters when a rank's launches are only 10-20 waves).
namespace {
struct SideStream {
    int dev = -1;
    cudaStream_t st = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    cudaError_t get(int device) {
        if (dev == device && st) return cudaSuccess;
        if (st) {
            cudaStreamDestroy(st);
            cudaEventDestroy(fork);
            cudaEventDestroy(join);
            st = nullptr;
        }
        cudaError_t e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&fork, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&join, cudaEventDisableTiming);
        dev = e == cudaSuccess ? device : -1;
        return e;
    }
};
thread_local SideStream g_side[16];  // one per device slot a thread drives
}  // namespace

// Diagonal bands: slabs [slab_begin, slab_end) of part `part` of `nparts` of the pair matrix
// (interleaved CTA groups of 8192 diagonals), thresholds refreshed per slab.  Parts (and, with a
// collective between calls, slab ranges) combine by element-wise minimum of d_best.
extern "C" int k4b_diag_slabs_device(k4b_packed *g, int both_strands, uint32_t part, uint32_t nparts,
                                     uint32_t slab_begin, uint32_t slab_end, uint32_t *d_best,
                                     void *stream, int *launches) {
    if (launches) *launches = 0;
    if (!g || !d_best) return fail(K4B_ERR_PARAMS, "NULL argument");
    if (!nparts || part >= nparts) return fail(K4B_ERR_PARAMS, "part %u of %u", part, nparts);
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t K = g->K, len = g->len;
    if (len < K) return K4B_OK;
    const bool three = g->has_non_acgt != 0, crick = both_strands != 0;
    const BandGeometry bg = band_geometry(g, crick, nparts);
    const uint32_t M = bg.M, n_slabs = bg.plan.n_slabs();
    slab_end = std::min(slab_end, n_slabs);
    if (slab_begin >= slab_end) return K4B_OK;
    int nl = 0;
    RC(diag_prepare(g, crick, st, &nl));
    const uint32_t bm_shift = 8;
    const uint32_t n_blocks = (M >> bm_shift) + 1;
    if (n_slabs > kMaxSlabs) return fail(K4B_ERR_PARAMS, "%u slabs exceed the %u bookkeeping slots", n_slabs, kMaxSlabs);
    DevScratch bm_buf;  // [n_blocks] block maxima + per slab: [kMaxSlabs] global maxima, [kMaxSlabs] low-block counts
    CU(bm_buf.alloc(((size_t)n_blocks + 2 * kMaxSlabs) * 4, st));
    uint32_t *d_bm = bm_buf.as<uint32_t>();
    CU(cudaMemsetAsync(d_bm + n_blocks, 0, 2 * kMaxSlabs * 4, st));
    const char *rs = getenv("K4B_DIAG_ROWS");
    const uint32_t rows_per_seg = rs ? (uint32_t)atoi(rs) : 8193u;  // 1 warm-up row + 256 whole 32-row blocks: no tail
    DiagParams dp;
    dp.a = g->view();
    dp.va = g->view();
    dp.vb = g->view();
    dp.K = K;
    dp.Mrow = dp.Mcol = M;
    dp.row_flip = 0;
    dp.update_cols = 1;
    dp.wild = 0;
    dp.t_fixed = 0;
    dp.rows_per_seg = rows_per_seg;
    dp.n_seg = (uint32_t)(((uint64_t)M + 1 + rows_per_seg - 1) / rows_per_seg);
    dp.best = d_best;
    dp.blockmax = d_bm;
    dp.bm_shift = bm_shift;
    dp.part = part;
    dp.nparts = nparts;
    dp.q_period = bg.plan.R;
    // counter width: one plane fewer whenever every threshold of the slab fits (checked on the
    // device against the slab's global maximum, so nothing waits for the host)
    const int np_full = diag_planes_for_k(K);
    const int np_small = (np_full - 1 >= 5 && !getenv("K4B_DIAG_FULLNP")) ? np_full - 1 : 0;
    // the narrow instance must lift thresholds below K+1-2^(np_small-1) (its counters cannot hold
    // K otherwise), which multiplies the flags there: it runs only while fewer than 1/64 of the
    // blocks sit below that floor (K=25: floor 10 but final minima 6-9 -> full width)
    const uint32_t low_floor = (np_small && K + 1 > (1u << (np_small - 1))) ? K + 1 - (1u << (np_small - 1)) : 0u;
    RC(g_tp.begin(g->device, slab_begin == 0, st));
    cudaError_t e = cudaSuccess;
    SideStream &side = g_side[g->device & 15];
    const bool fork = crick && !getenv("K4B_DIAG_ONE_STREAM");
    if (fork) e = side.get(g->device);
    for (uint32_t slab = slab_begin; slab < slab_end && e == cudaSuccess; ++slab) {
        uint32_t *d_tmax = d_bm + n_blocks + slab, *d_low = d_bm + n_blocks + kMaxSlabs + slab;
        e = launch_blockmax(d_best, g->view(), M + 1, bm_shift, d_bm, n_blocks, d_tmax, low_floor, d_low, st);
        if (fork && e == cudaSuccess) e = cudaEventRecord(side.fork, st);
        if (fork && e == cudaSuccess) e = cudaStreamWaitEvent(side.st, side.fork, 0);
        dp.tmax_ptr = d_tmax;
        dp.sel_limit = np_small ? (1u << (np_small - 1)) : 0u;
        dp.low_ptr = d_low;
        dp.low_max = n_blocks / 64;
        ++nl;
        dp.q_lo = bg.plan.bounds[slab];
        dp.q_span = bg.plan.bounds[slab + 1] - dp.q_lo;
        for (int strand = 0; strand < (crick ? 2 : 1) && e == cudaSuccess; ++strand) {
            const uint64_t ngroups_all = strand ? bg.gc : bg.gw;
            if (part >= ngroups_all) continue;
            const uint64_t Qs = (ngroups_all - part + nparts - 1) / nparts;  // this part's groups
            const uint64_t rem = Qs % bg.plan.R;
            const uint64_t ng = Qs / bg.plan.R * dp.q_span +
                                std::min<uint64_t>(dp.q_span, rem > dp.q_lo ? rem - dp.q_lo : 0);
            dp.mode = strand ? kDiagCrick : kDiagWatson;
            dp.col_flip = strand;
            dp.b = strand ? g->rc_view() : g->view();
            dp.s_first = strand ? -(long long)M : 1;
            DiagParams q = dp;
            // launches of fewer than ~40 waves of CTAs (small inputs, many parts): half-length row segments,
            // i.e. twice as many CTAs of half the duration - a shorter tail for 0.6 % more warm-up rows
            if (!rs && ng * dp.n_seg < 47360ull) {
                q.rows_per_seg = 4097u;
                q.n_seg = (uint32_t)(((uint64_t)M + 1 + q.rows_per_seg - 1) / q.rows_per_seg);
            }
            // the 1-D grid is limited to 2^31-1 CTAs: split very large launches by groups
            const uint32_t max_groups = std::max(1u, 0x7fffffffu / q.n_seg);
            cudaStream_t ls = (fork && strand) ? side.st : st;  // Crick launches overlap the Watson ones
            for (uint64_t done = 0; done < ng && e == cudaSuccess; done += max_groups) {
                q.l_first = (uint32_t)done;
                const uint32_t now = (uint32_t)std::min<uint64_t>(max_groups, ng - done);
                if (np_small) {
                    q.sel = 1;
                    e = launch_diag(q, three, np_small, now, ls, nullptr);
                    ++nl;
                    q.sel = 2;
                } else {
                    q.sel = 0;
                }
                if (e == cudaSuccess) e = launch_diag(q, three, np_full, now, ls, nullptr);
                ++nl;
            }
        }
        if (fork && e == cudaSuccess) e = cudaEventRecord(side.join, side.st);
        if (fork && e == cudaSuccess) e = cudaStreamWaitEvent(st, side.join, 0);
    }
    if (e == cudaSuccess) e = g_tp.end(st);
    if (e == cudaSuccess && !g_h_tmax) e = cudaMallocHost(&g_h_tmax, 2 * kMaxSlabs * sizeof(uint32_t));
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(g_h_tmax + slab_begin, d_bm + n_blocks + slab_begin,
                            (slab_end - slab_begin) * sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(g_h_tmax + kMaxSlabs + slab_begin, d_bm + n_blocks + kMaxSlabs + slab_begin,
                            (slab_end - slab_begin) * sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
    g_info_low_max = n_blocks / 64;
    g_info_slabs = n_slabs;
    g_info_np_full = (uint32_t)np_full;
    g_info_np_small = (uint32_t)np_small;
    g_info_limit = np_small ? (1u << (np_small - 1)) : 0u;
    if (e != cudaSuccess) return fail(cuda_code(e), "diagonal engine launch: %s", cudaGetErrorString(e));
    if (launches) *launches = nl;
    return K4B_OK;
}

// all slabs of one part in one call
extern "C" int k4b_diag_bands_device(k4b_packed *g, int both_strands, uint32_t part, uint32_t nparts,
                                     uint32_t *d_best, void *stream, int *launches) {
    return k4b_diag_slabs_device(g, both_strands, part, nparts, 0, 0xffffffffu, d_best, stream, launches);
}

// P(distance of two unrelated K-mers < t) for uniform random bases: the binomial(K, 3/4) tail
static double random_pair_tail(uint32_t K, uint32_t t) {
    double sum = 0;
    for (uint32_t x = 0; x < t && x <= K; ++x)
        sum += exp(lgamma(K + 1.0) - lgamma(x + 1.0) - lgamma(K - x + 1.0) + x * log(0.75) + (K - x) * log(0.25));
    return sum;
}

// Targeted (probes vs assembly) on the band engine: rows = probe K-mers (forward, then their
// reverse complement), columns = target K-mers, fixed threshold = the "not found" clamp, rows
// only.  Part `part` of `nparts` of the diagonals; parts combine by element-wise minimum.
extern "C" int k4b_targeted_diag_device(k4b_packed *probes, k4b_packed *targets, int both_strands,
                                        uint32_t clamp, uint32_t part, uint32_t nparts,
                                        uint32_t *d_best, void *stream, int *launches) {
    if (launches) *launches = 0;
    if (!probes || !targets || !d_best) return fail(K4B_ERR_PARAMS, "NULL argument");
    if (probes->K != targets->K) return fail(K4B_ERR_PARAMS, "probe/target K differ");
    if (probes->device != targets->device) return fail(K4B_ERR_PARAMS, "images live on different devices");
    if (!nparts || part >= nparts || !clamp) return fail(K4B_ERR_PARAMS, "bad part/clamp");
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t K = probes->K;
    if (probes->len < K || targets->len < K) return K4B_OK;
    const bool three = probes->has_non_acgt || targets->has_non_acgt, crick = both_strands != 0;
    int nl = 0;
    RC(diag_prepare(probes, crick, st, &nl));
    const char *rs = getenv("K4B_DIAG_ROWS");
    DiagParams dp;
    dp.b = targets->view();
    dp.va = probes->view();
    dp.vb = targets->view();
    dp.K = K;
    dp.Mrow = probes->len - K;
    dp.Mcol = targets->len - K;
    dp.mode = kDiagRect;
    dp.col_flip = 0;
    dp.update_cols = 0;
    dp.wild = three ? 1 : 0;
    dp.t_fixed = std::min(clamp, K + 1);
    dp.rows_per_seg = rs ? (uint32_t)atoi(rs) : 8193u;
    dp.n_seg = (uint32_t)(((uint64_t)dp.Mrow + 1 + dp.rows_per_seg - 1) / dp.rows_per_seg);
    dp.best = d_best;
    dp.blockmax = nullptr;
    dp.bm_shift = 0;
    dp.tmax_ptr = nullptr;
    dp.low_ptr = nullptr;
    dp.low_max = 0;
    dp.sel = 0;
    dp.sel_limit = 0;
    dp.s_first = -(long long)dp.Mrow;
    dp.part = part;
    dp.nparts = nparts;
    dp.q_period = 1;  // no slabs: the threshold is fixed
    dp.q_lo = 0;
    dp.q_span = 1;
    int np = diag_planes_for_k(K);
    // the clamp is tiny: one counter plane fewer, provided the floor the narrow counters impose
    // on the threshold (K+1-2^(np-2)) does not flag too many cells of unrelated sequence
    if (np - 1 >= 5 && dp.t_fixed <= (1u << (np - 2))) {
        const uint32_t half = 1u << (np - 2);
        const uint32_t floor_narrow = K + 1 > half ? K + 1 - half : 0u;
        if (floor_narrow <= dp.t_fixed || random_pair_tail(K, floor_narrow) < 1e-7) --np;
    }
    const uint64_t groups = ((uint64_t)dp.Mrow + dp.Mcol + 1 + kDiagGroupDiagonals - 1) / kDiagGroupDiagonals;
    RC(g_tp.begin(probes->device, true, st));
    cudaError_t e = cudaSuccess;
    for (int strand = 0; strand < (crick ? 2 : 1) && e == cudaSuccess; ++strand) {
        if (part >= groups) break;
        dp.a = strand ? probes->rc_view() : probes->view();
        dp.row_flip = strand;
        const uint32_t ng = (uint32_t)((groups - part + nparts - 1) / nparts);
        const uint32_t max_groups = std::max(1u, 0x7fffffffu / dp.n_seg);
        for (uint32_t done = 0; done < ng && e == cudaSuccess; done += max_groups) {
            dp.l_first = done;
            e = launch_diag(dp, three, np, std::min(max_groups, ng - done), st, nullptr);
            ++nl;
        }
    }
    if (e == cudaSuccess) e = g_tp.end(st);
    if (e != cudaSuccess) return fail(cuda_code(e), "targeted band launch: %s", cudaGetErrorString(e));
    if (launches) *launches = nl;
    return K4B_OK;
}

// statistics of the most recent seed-engine call of this thread (pinned, copied on its stream)
static thread_local unsigned long long *g_h_seed = nullptr;
extern "C" int k4b_last_seed_info(uint64_t *occurrences, uint64_t *indexed_cores) {
    if (!g_h_seed || !g_tp.used) return fail(K4B_ERR_PARAMS, "no seed-engine call recorded on this thread");
    CU(cudaDeviceSynchronize());
    uint64_t occ = 0;
    for (int i = 0; i < kSeedOccSlots; ++i) occ += g_h_seed[i];
    if (occurrences) *occurrences = occ;
    if (indexed_cores) *indexed_cores = g_h_seed[kSeedOccSlots] & 0xffffffffull;
    return K4B_OK;
}

// Targeted (probes vs assembly) by seed-and-verify (k4b_seed.cu): exact for every distance below
// clamp, the "not found" value, as long as clamp <= K / core_len (pigeonhole over the disjoint
// cores).  Probe K-mers holding N / InDel are skipped (d_best keeps its value): wildcard probes need
// the reference's substitution rule and belong to the brute-force engines.  probes == targets (the
// same handle) selects the rules of probes drawn from the assembly itself.  Two ways to split the
// work, both combining by element-wise minimum of d_best: a RANGE of probe K-mers [q_begin, q_end)
// against the whole index, or (part, nparts) = a share of the index BUCKETS against all probes -
// then only 1/nparts of the index is built, which is what scales when the index build matters.
// Where the reference's answer may differ from the exact one.  CSfxArray::LocateHamming walks at most
// MaxCoreDepth suffix-array entries per core and gives a core up after MinCoreDepth entries when it has
// more copies than that (SfxArray.cpp:4480-4494); hammings passes MaxCoreDepth = {10000, 20000, 50000,
// 5000} for -s 0..3 (hammings.cpp:2366-2386), multiplied by 10 / 8 / 4 / 2 / 1 on the cascade levels
// Rmax = 1 / 2 / 3 / 4 / >= 5 (SfxArray.cpp:4297-4311).  The seed engine has no such cut; it can flag the
// probe K-mers that hold a core (of the last cascade level = the index's core length) with more copies.
static thread_local struct {
    int sensitivity = 0;          // -s of the host (k4b_set_reference_sensitivity)
    uint32_t cap = 0;             // device-API watch: entries per bucket above which a probe is flagged
    uint8_t *d_deep = nullptr;    // device-API watch: one byte per probe position (caller-owned)
    unsigned long long last_count = 0;  // host-buffer calls: flagged probe K-mers answered below "not found"
    uint32_t last_cap = 0;
    bool last_valid = false;
} g_depth;

static uint32_t reference_depth_cap(int sensitivity, int R) {
    static const uint32_t max_depth[4] = {10000u, 20000u, 50000u, 5000u};
    const uint32_t mult = R <= 1 ? 10u : R == 2 ? 8u : R == 3 ? 4u : R == 4 ? 2u : 1u;
    return mult * max_depth[sensitivity & 3];
}

extern "C" int k4b_set_reference_sensitivity(int sensitivity) {
    if (sensitivity < 0 || sensitivity > 3) return fail(K4B_ERR_PARAMS, "sensitivity %d outside 0..3", sensitivity);
    g_depth.sensitivity = sensitivity;
    return K4B_OK;
}

extern "C" int k4b_seed_watch_depth(uint32_t max_copies, uint8_t *d_deep) {
    g_depth.cap = max_copies;
    g_depth.d_deep = d_deep;
    return K4B_OK;
}

extern "C" int k4b_last_depth_cut(uint64_t *probe_kmers, uint32_t *max_copies) {
    if (!g_depth.last_valid) return fail(K4B_ERR_PARAMS, "no seed-engine host call recorded on this thread");
    if (probe_kmers) *probe_kmers = g_depth.last_count;
    if (max_copies) *max_copies = g_depth.last_cap;
    return K4B_OK;
}

static int seed_run(k4b_packed *probes, k4b_packed *targets, int both_strands, uint32_t clamp, uint32_t core_len,
                    uint32_t q_begin, uint32_t q_end, uint32_t part, uint32_t nparts, uint32_t *d_best, void *stream,
                    int *launches) {
    if (launches) *launches = 0;
    if (!probes || !targets || !d_best) return fail(K4B_ERR_PARAMS, "NULL argument");
    if (probes->K != targets->K) return fail(K4B_ERR_PARAMS, "probe/target K differ");
    if (probes->device != targets->device) return fail(K4B_ERR_PARAMS, "images live on different devices");
    if (!nparts || part >= nparts) return fail(K4B_ERR_PARAMS, "part %u of %u", part, nparts);
    const uint32_t K = probes->K;
    if (!core_len || core_len > K || !clamp || clamp > K / core_len)
        return fail(K4B_ERR_PARAMS, "core_len=%u clamp=%u: the pigeonhole bound needs clamp <= K/core_len", core_len, clamp);
    cudaStream_t st = (cudaStream_t)stream;
    if (q_end > probes->len) q_end = probes->len;
    if (probes->len < K || targets->len < K || q_begin >= q_end) return K4B_OK;
    CU(cudaSetDevice(probes->device));
    const bool crick = both_strands != 0;
    int nl = 0;
    RC(diag_prepare(probes, crick, st, &nl));  // reverse-complemented probe planes
    const uint32_t nb = 1u << seed_bucket_bits(core_len);
    const uint32_t b_lo = (uint32_t)((uint64_t)nb * part / nparts), b_hi = (uint32_t)((uint64_t)nb * (part + 1) / nparts);
    const size_t temp_bytes = seed_scan_temp_bytes(nb);
    // index: 16-byte entries | occ slots | cnt | off | cursor (nb+1 each); scan scratch; -z entry starts
    DevScratch idx_buf, tmp_buf, ent_buf;
    const size_t ent_bytes = (size_t)targets->len * sizeof(uint4);
    CU(idx_buf.alloc(ent_bytes + (size_t)kSeedOccSlots * 8 + 3 * ((size_t)nb + 1) * 4, st));
    CU(tmp_buf.alloc(temp_bytes, st));
    uint4 *d_ent = idx_buf.as<uint4>();
    unsigned long long *d_occ = (unsigned long long *)((char *)idx_buf.p + ent_bytes);
    uint32_t *d_cnt = (uint32_t *)(d_occ + kSeedOccSlots), *d_off = d_cnt + nb + 1, *d_cur = d_off + nb + 1;
    CU(cudaMemsetAsync(d_occ, 0, (size_t)kSeedOccSlots * 8 + ((size_t)nb + 1) * 4, st));  // occ slots + counters
    // probes == targets: the K-mers of the assembly against the assembly itself (exact sense self hit
    // skipped; -z of the k4b_hamm_targeted_z call in progress on this thread)
    SeedSelfRules self{probes == targets ? 1 : 0, 0, nullptr, 0};
    if (self.on && g_zf.mode && !g_zf.starts.empty()) {
        CU(ent_buf.alloc(g_zf.starts.size() * 4, st));
        CU(cudaMemcpyAsync(ent_buf.p, g_zf.starts.data(), g_zf.starts.size() * 4, cudaMemcpyHostToDevice, st));
        self.zfilt = g_zf.mode;
        self.ent_starts = ent_buf.as<uint32_t>();
        self.n_ent = (uint32_t)g_zf.starts.size();
    }
    RC(g_tp.begin(probes->device, true, st));
    int nq = 0, ni = 0;
    // index build: entries placed one by one, or (default; K4B_SEED_INDEX=0 turns it off; buckets of <= 16 bits) by two partition
    // passes through a scratch array as large as the index; without room for it the first build is used
    DevScratch part_buf;
    const char *ix = getenv("K4B_SEED_INDEX");
    const int part_mode = ix ? atoi(ix) : kSeedIndexDefault;
    if (part_mode != 0 && seed_index_can_partition(core_len) && part_buf.alloc(ent_bytes, st) != cudaSuccess) {
        part_buf.p = nullptr;
        cudaGetLastError();  // out of memory is no error here
    }
    cudaError_t e = launch_seed_index(targets->view(), core_len, b_lo, b_hi, d_cnt, d_off, d_cur, d_ent,
                                      part_buf.as<uint4>(), part_mode, tmp_buf.p, temp_bytes, st, &ni);
    part_buf.release();  // stream-ordered: the join's sort buffers may take its place
    if (e == cudaSuccess)
        e = launch_seed_query(probes->view(), crick ? probes->rc_view() : probes->view(), targets->view(), K, core_len,
                              d_off, d_ent, q_begin, q_end, b_lo, b_hi, clamp, crick, targets->has_non_acgt != 0,
                              probes->has_non_acgt != 0, self, g_depth.cap, g_depth.d_deep, d_best, d_occ, st, &nq);
    if (e == cudaSuccess) e = g_tp.end(st);
    if (e == cudaSuccess && !g_h_seed) e = cudaMallocHost(&g_h_seed, (kSeedOccSlots + 1) * 8);
    if (e == cudaSuccess) e = cudaMemcpyAsync(g_h_seed, d_occ, kSeedOccSlots * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess)  // number of indexed cores = the last bucket offset
        e = cudaMemcpyAsync(g_h_seed + kSeedOccSlots, d_off + nb, 4, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return fail(cuda_code(e), "seed engine launch: %s", cudaGetErrorString(e));
    if (launches) *launches = nl + ni + nq;  // reverse complement of the probes, the index passes, the query kernels
    return K4B_OK;
}

extern "C" int k4b_targeted_seed_device(k4b_packed *probes, k4b_packed *targets, int both_strands,
                                        uint32_t clamp, uint32_t core_len, uint32_t q_begin,
                                        uint32_t q_end, uint32_t *d_best, void *stream, int *launches) {
    return seed_run(probes, targets, both_strands, clamp, core_len, q_begin, q_end, 0, 1, d_best, stream, launches);
}

extern "C" int k4b_targeted_seed_part_device(k4b_packed *probes, k4b_packed *targets, int both_strands,
                                             uint32_t clamp, uint32_t core_len, uint32_t part, uint32_t nparts,
                                             uint32_t *d_best, void *stream, int *launches) {
    return seed_run(probes, targets, both_strands, clamp, core_len, 0, 0xffffffffu, part, nparts, d_best, stream,
                    launches);
}

// minima -> uint16 with the targeted rules applied (cap at clamp, > 4 wildcards -> 0)
extern "C" int k4b_targeted_finalize_device(k4b_packed *probes, const uint32_t *d_best, uint32_t clamp,
                                            uint16_t *d_out_min, void *stream) {
    if (!probes || !d_best || !d_out_min) return fail(K4B_ERR_PARAMS, "NULL argument");
    CU(cudaSetDevice(probes->device));
    CU(launch_finalize(d_best, probes->view(), 0, probes->len, probes->K, clamp, 4, d_out_min, (cudaStream_t)stream));
    return K4B_OK;
}

// bootstrap of every K-mer + this part's bands (single-call form)
extern "C" int k4b_exhaustive_diag_device(k4b_packed *g, int both_strands, uint32_t part,
                                          uint32_t nparts, uint32_t *d_best, void *stream,
                                          int *launches) {
    if (launches) *launches = 0;
    if (!g) return fail(K4B_ERR_PARAMS, "NULL argument");
    RC(k4b_diag_bootstrap_device(g, both_strands, 0, g->len, d_best, stream));
    int nl = 0;
    RC(k4b_diag_bands_device(g, both_strands, part, nparts, d_best, stream, &nl));
    if (launches) *launches = nl + 1;
    return K4B_OK;
}

// ------------------------------------------------------------------------------------------
// host-buffer entry points
// ------------------------------------------------------------------------------------------
namespace {
struct DevJob {
    k4b_packed *q = nullptr, *t = nullptr;  // t == q for all-vs-all
    DevScratch d_out;                       // stream-ordered, back to the pool when the job dies
    uint16_t *h_out = nullptr;              // the engine's cached pinned staging of this device slot
    uint32_t q_begin = 0, q_end = 0;
    void release() {
        d_out.release();
        if (t && t != q) k4b_packed_free(t);
        if (q) k4b_packed_free(q);
        q = t = nullptr;
    }
};

// broadcast the image of `src` (on device slot 0 of the engine) to fresh pooled images on every
// other engine device: ONE ncclBroadcast per set, nothing else crosses NVLink afterwards
int broadcast_packed(k4b_packed *src, std::vector<k4b_packed *> &out) {
    const int n = (int)g_eng.devs.size();
    out.assign(n, nullptr);
    out[0] = src;
    if (n == 1) return 0;
    const size_t bytes = k4b_packed_image_size(src);
    for (int i = 1; i < n; ++i) {
        CU(cudaSetDevice(g_eng.devs[i]));
        void *img = nullptr;
        CU(cudaMallocAsync(&img, bytes, g_eng.streams[i]));
        int rc = k4b_packed_from_image(img, bytes, src->len, src->K, src->has_non_acgt, &out[i]);
        if (rc) {
            cudaFreeAsync(img, g_eng.streams[i]);
            return rc;
        }
        out[i]->owns = true;
        out[i]->pooled = true;
        out[i]->pool_stream = g_eng.streams[i];
        out[i]->num_kmers = src->num_kmers;
    }
    int nr = g_eng.nccl.GroupStart();
    for (int i = 0; i < n && !nr; ++i)
        nr = g_eng.nccl.Broadcast(src->d_image, out[i]->d_image, bytes, kNcclUint8, 0,
                                  g_eng.comms[i], g_eng.streams[i]);
    int nr2 = g_eng.nccl.GroupEnd();
    if (nr || nr2)
        return fail(K4B_ERR_NCCL, "ncclBroadcast: %s", g_eng.nccl.GetErrorString(nr ? nr : nr2));
    return 0;  // stream ordered: every later use of the images is enqueued on the same streams
}

// K4B_TRACE=1 prints the host-side phase times of each host-buffer call to stderr
struct PhaseTrace {
    bool on = getenv("K4B_TRACE") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now(), last = t0;
    void mark(const char *what) {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[k4b trace] %-22s %9.3f ms (total %9.3f ms)\n", what,
                std::chrono::duration<double, std::milli>(now - last).count(),
                std::chrono::duration<double, std::milli>(now - t0).count());
        last = now;
    }
};

// queries [q_begin,q_end) split evenly by position over the engine's devices; per-device
// minima land in pinned host buffers and are handed to `sink(pos, value)`
template <typename Sink>
int run_sharded(const uint8_t *q_concat, uint32_t q_len, const uint8_t *t_concat, uint32_t t_len,
                uint32_t K, int both, int self_ex, uint32_t q_begin, uint32_t q_end,
                uint32_t clamp, int use_all_devices, const SweepRange &sweep, Sink sink) {
    RC(ensure_init());
    const int n = use_all_devices ? (int)g_eng.devs.size() : 1;
    if (q_end > q_len) q_end = q_len;
    if (q_begin >= q_end) return K4B_OK;
    std::vector<DevJob> jobs(n);
    PhaseTrace trace;
    auto cleanup = [&]() {
        for (int i = 0; i < n; ++i) {
            cudaSetDevice(g_eng.devs[i]);
            jobs[i].release();
        }
        cudaSetDevice(g_eng.devs[0]);
    };
    int rc = 0;
    do {
        // pack once on device 0, broadcast to the rest
        CU(cudaSetDevice(g_eng.devs[0]));
        k4b_packed *q0 = nullptr, *t0 = nullptr;
        if ((rc = pack_host_pooled(q_concat, q_len, K, 0, &q0))) break;
        jobs[0].q = q0;
        const bool same = (t_concat == nullptr);
        if (same) {
            jobs[0].t = q0;
        } else {
            if ((rc = pack_host_pooled(t_concat, t_len, K, 0, &t0))) break;
            jobs[0].t = t0;
        }
        if (n > 1) {
            std::vector<k4b_packed *> qs, ts;
            rc = broadcast_packed(q0, qs);
            for (int i = 1; i < n; ++i) jobs[i].q = qs.size() > (size_t)i ? qs[i] : nullptr;
            if (rc) break;
            if (same) {
                for (int i = 1; i < n; ++i) jobs[i].t = jobs[i].q;
            } else {
                rc = broadcast_packed(t0, ts);
                for (int i = 1; i < n; ++i) jobs[i].t = ts.size() > (size_t)i ? ts[i] : nullptr;
                if (rc) break;
            }
        }
        trace.mark("H2D + pack (+bcast)");
        // launch every shard (asynchronous), then collect
        const uint64_t span = (uint64_t)q_end - q_begin;
        for (int i = 0; i < n && !rc; ++i) {
            DevJob &j = jobs[i];
            j.q_begin = q_begin + (uint32_t)(span * i / n);
            j.q_end = q_begin + (uint32_t)(span * (i + 1) / n);
            const uint32_t nq = j.q_end - j.q_begin;
            if (!nq) continue;
            cudaError_t e = cudaSetDevice(g_eng.devs[i]);
            if (e == cudaSuccess) e = j.d_out.alloc((size_t)nq * 2, g_eng.streams[i]);
            if (e == cudaSuccess && !(j.h_out = (uint16_t *)pinned_result(i, (size_t)nq * 2))) e = cudaErrorMemoryAllocation;
            if (e != cudaSuccess) {
                rc = fail(cuda_code(e), "shard buffers: %s", cudaGetErrorString(e));
                break;
            }
            rc = allpairs_impl(j.q, j.t, both, self_ex, j.q_begin, j.q_end, clamp, sweep, j.d_out.as<uint16_t>(),
                               g_eng.streams[i], nullptr);
            if (rc) break;
            e = cudaMemcpyAsync(j.h_out, j.d_out.p, (size_t)nq * 2, cudaMemcpyDeviceToHost,
                                g_eng.streams[i]);
            if (e != cudaSuccess) rc = fail(cuda_code(e), "D2H: %s", cudaGetErrorString(e));
        }
        trace.mark("alloc + launch");
        for (int i = 0; i < n; ++i) {
            cudaSetDevice(g_eng.devs[i]);
            cudaError_t e = cudaStreamSynchronize(g_eng.streams[i]);
            if (e != cudaSuccess && !rc)
                rc = fail(cuda_code(e), "device %d: %s", g_eng.devs[i], cudaGetErrorString(e));
        }
        if (rc) break;
        trace.mark("kernels + D2H");
        // concatenate per-GPU minima back on the host
        for (int i = 0; i < n; ++i) {
            const DevJob &j = jobs[i];
            for (uint32_t p = j.q_begin; p < j.q_end; ++p) sink(p, j.h_out[p - j.q_begin]);
        }
        trace.mark("host gather");
    } while (0);
    cleanup();
    trace.mark("free");
    return rc;
}

// Pair sets selected by sweep instances SSeqStart..SSeqEnd (hammings.cpp:883-939): Watson
// offset s for SSeqStart <= s <= min(SSeqEnd, len-K); Crick offset s-1, which walks the
// anti-diagonal c = len-1-(s-1) and, for s-1 >= 1, after wrapping also c = 2len-K-(s-1)
// (hammings.cpp:3345-3347, :3410-3415).  The full range selects every pair.
int make_sweep(uint32_t len, uint32_t K, uint32_t sweep_start, uint32_t sweep_end, SweepRange &r) {
    r = SweepRange();
    if (sweep_start == 0) return fail(K4B_ERR_PARAMS, "sweep start must be >= 1");
    const long long L = len, k = K;
    long long se = sweep_end == 0 ? L + 2 : (long long)sweep_end;
    const long long ss = sweep_start;
    const bool full = ss == 1 && se >= L + 1 - k;
    if (full || L < k) return 0;
    r.ranged = true;
    r.w_lo = ss;
    r.w_hi = std::min(se, L - k);
    const long long sp_lo = ss - 1, sp_hi = std::min(se, L + 1 - k) - 1;
    if (sp_hi >= sp_lo) {
        r.c1_lo = L - 1 - sp_hi;
        r.c1_hi = L - 1 - sp_lo;
        const long long lo2 = std::max(sp_lo, 1LL);
        if (sp_hi >= lo2) {
            r.c2_lo = 2 * L - k - sp_hi;
            r.c2_hi = 2 * L - k - lo2;
        }
    }
    return 0;
}
}  // namespace

// full-sweep exhaustive run on the diagonal engine: the pair matrix (not the queries) is
// partitioned over the devices, each keeps a complete array of running minima, and the arrays
// meet in ncclAllReduce(min) after the bootstrap and after every slab - the exchange step of
// the symmetric formulation.  All device memory comes from the stream-ordered pool, the result
// staging is the engine's cached pinned buffer: a warm call allocates nothing from the driver.
static int run_exhaustive_diag(const uint8_t *concat, uint32_t len, uint32_t K, int both,
                               uint16_t *out_min) {
    RC(ensure_init());
    const int n = (int)g_eng.devs.size();
    PhaseTrace trace;
    std::vector<k4b_packed *> imgs;
    std::vector<DevScratch> bests(n);
    DevScratch d_out;
    int rc = 0;
    do {
        k4b_packed *g0 = nullptr;
        if ((rc = pack_host_pooled(concat, len, K, 0, &g0))) break;
        rc = broadcast_packed(g0, imgs);
        if (imgs.empty()) imgs.push_back(g0);
        if (rc) break;
        trace.mark("H2D + pack (+bcast)");
        for (int i = 0; i < n && !rc; ++i) {
            cudaError_t e = cudaSetDevice(g_eng.devs[i]);
            if (e == cudaSuccess) e = bests[i].alloc((size_t)len * 4, g_eng.streams[i]);
            if (e != cudaSuccess) {
                rc = fail(cuda_code(e), "minima buffer: %s", cudaGetErrorString(e));
                break;
            }
            rc = k4b_best_init_device(bests[i].as<uint32_t>(), len, K, g_eng.streams[i]);
            // bootstrap: every device takes a shard of the queries
            const uint32_t qb = (uint32_t)((uint64_t)len * i / n), qe = (uint32_t)((uint64_t)len * (i + 1) / n);
            if (!rc) rc = k4b_diag_bootstrap_device(imgs[i], both, qb, qe, bests[i].as<uint32_t>(), g_eng.streams[i]);
        }
        if (rc) break;
        auto allreduce_min = [&]() -> int {
            if (n == 1) return 0;
            int nr = g_eng.nccl.GroupStart();
            for (int i = 0; i < n && !nr; ++i)
                nr = g_eng.nccl.AllReduce(bests[i].p, bests[i].p, len, kNcclUint32, kNcclMin, g_eng.comms[i], g_eng.streams[i]);
            const int nr2 = g_eng.nccl.GroupEnd();
            if (nr || nr2) return fail(K4B_ERR_NCCL, "ncclAllReduce(min): %s", g_eng.nccl.GetErrorString(nr ? nr : nr2));
            return 0;
        };
        if ((rc = allreduce_min())) break;
        // bands slab by slab; with several devices the minima meet after EVERY slab, so that each
        // device thresholds against what all of them have found so far
        uint32_t n_slabs = 0;
        if ((rc = k4b_diag_slab_count(imgs[0], both, (uint32_t)n, &n_slabs))) break;
        for (uint32_t slab = 0; slab < n_slabs && !rc; slab = (n == 1 ? n_slabs : slab + 1)) {
            for (int i = 0; i < n && !rc; ++i) {
                cudaSetDevice(g_eng.devs[i]);
                rc = k4b_diag_slabs_device(imgs[i], both, (uint32_t)i, (uint32_t)n, slab,
                                           n == 1 ? n_slabs : slab + 1, bests[i].as<uint32_t>(), g_eng.streams[i], nullptr);
            }
            if (!rc) rc = allreduce_min();
        }
        if (rc) break;
        trace.mark("launch");
        cudaError_t e = cudaSetDevice(g_eng.devs[0]);
        uint16_t *h_out = nullptr;
        if (e == cudaSuccess) e = d_out.alloc((size_t)len * 2, g_eng.streams[0]);
        if (e == cudaSuccess && !(h_out = (uint16_t *)pinned_result(0, (size_t)len * 2))) e = cudaErrorMemoryAllocation;
        if (e != cudaSuccess) {
            rc = fail(cuda_code(e), "result buffers: %s", cudaGetErrorString(e));
            break;
        }
        if ((rc = k4b_best_finalize_device(imgs[0], bests[0].as<uint32_t>(), d_out.as<uint16_t>(), g_eng.streams[0]))) break;
        e = cudaMemcpyAsync(h_out, d_out.p, (size_t)len * 2, cudaMemcpyDeviceToHost, g_eng.streams[0]);
        for (int i = 0; i < n; ++i) {
            cudaSetDevice(g_eng.devs[i]);
            const cudaError_t e2 = cudaStreamSynchronize(g_eng.streams[i]);
            if (e2 != cudaSuccess && e == cudaSuccess) e = e2;
        }
        if (e != cudaSuccess) {
            rc = fail(cuda_code(e), "diagonal engine: %s", cudaGetErrorString(e));
            break;
        }
        trace.mark("kernels + D2H");
        for (uint32_t p = 0; p < len; ++p)
            if (h_out[p] <= K && h_out[p] < out_min[p]) out_min[p] = h_out[p];
        trace.mark("host gather");
    } while (0);
    for (int i = 0; i < n; ++i) {
        cudaSetDevice(g_eng.devs[i]);
        bests[i].release();
        if ((size_t)i < imgs.size() && imgs[i]) k4b_packed_free(imgs[i]);
    }
    cudaSetDevice(g_eng.devs[0]);
    d_out.release();
    trace.mark("free");
    return rc;
}

// Whole-probe-set targeted run on the seed-and-verify engine (pure-ACGT probes, cores of at
// least 6 bases) or on the band engine; *used = 0 when neither applies (caller falls back to the
// POPC engine).  Devices split the index BUCKETS (seed: each builds and joins its share of the
// index against all probes) or the diagonals (bands); minima meet in one ncclAllReduce(min).
static int run_targeted_big(const uint8_t *t_concat, uint32_t t_len, const uint8_t *q_concat, uint32_t q_len,
                            uint32_t K, int R, int both, uint32_t clamp, uint32_t core_len, bool allow_seed,
                            bool allow_diag, uint8_t *out_h, int *used) {
    *used = 0;
    RC(ensure_init());
    const int n = (int)g_eng.devs.size();
    PhaseTrace trace;
    std::vector<k4b_packed *> qs, ts;
    std::vector<DevScratch> bests(n), deeps(n);
    DevScratch d_out, d_deep_count;
    const uint32_t depth_cap = reference_depth_cap(g_depth.sensitivity, R);
    g_depth.last_valid = false;
    bool use_seed = false;
    int rc = 0;
    do {
        k4b_packed *q0 = nullptr, *t0 = nullptr;
        const bool same = q_concat == nullptr;  // probes = the K-mers of the assembly itself (no -I)
        if (same) {
            q_concat = t_concat;
            q_len = t_len;
        }
        if ((rc = pack_host_pooled(q_concat, q_len, K, 0, &q0))) break;
        qs.push_back(q0);
        // a core of c bases occurs in ~len/4^c places: below 6 bases verifying every occurrence costs
        // as much as the bit-sliced bands
        bool seed = allow_seed && core_len >= 6 && clamp <= K / core_len;
        // probe K-mers that hold N / InDel (wildcards, SfxArray.cpp:4266-4296) cannot be seeded: their
        // positions are collected as intervals and answered by the POPC engine afterwards
        std::vector<std::pair<uint32_t, uint32_t>> impure;
        if (seed && q0->has_non_acgt) {
            uint64_t total = 0;
            uint32_t in_win = 0, run_b = 0;
            bool open = false;
            for (uint32_t p = 0; p < K - 1 && p < q_len; ++p) in_win += (q_concat[p] >= 4 && q_concat[p] < 7);
            for (uint32_t p = 0; p + K <= q_len; ++p) {
                in_win += (q_concat[p + K - 1] >= 4 && q_concat[p + K - 1] < 7);
                if (in_win && !open) { open = true; run_b = p; }
                if (!in_win && open) { open = false; impure.emplace_back(run_b, p); total += p - run_b; }
                in_win -= (q_concat[p] >= 4 && q_concat[p] < 7);
            }
            if (open) { impure.emplace_back(run_b, q_len - K + 1); total += q_len - K + 1 - run_b; }
            if (impure.size() > 4096 || total * 4 > q_len) seed = false;  // mostly wildcards: not worth it
        }
        if (same && !seed) break;          // the band engine has no self-hit rules
        if (!seed && !allow_diag) break;  // *used stays 0: the caller runs the POPC engine
        *used = 1;
        if (same) {
            t0 = q0;
        } else if ((rc = pack_host_pooled(t_concat, t_len, K, 0, &t0))) {
            break;
        }
        ts.push_back(t0);
        use_seed = seed;
        if (n > 1) {
            qs.clear();
            ts.clear();
            rc = broadcast_packed(q0, qs);
            if (!rc && same) ts = qs;
            if (!rc && !same) rc = broadcast_packed(t0, ts);
            if (rc) {
                if (qs.empty()) qs.push_back(q0);
                if (ts.empty()) ts.push_back(t0);
                break;
            }
        }
        trace.mark("H2D + pack (+bcast)");
        uint16_t *h_out = nullptr;
        {
            cudaError_t e = cudaSetDevice(g_eng.devs[0]);
            if (e == cudaSuccess) e = d_out.alloc((size_t)q_len * 2, g_eng.streams[0]);
            // results + (after them, 8-byte aligned) the depth-watch count
            if (e == cudaSuccess && !(h_out = (uint16_t *)pinned_result(0, (size_t)q_len * 2 + 16))) e = cudaErrorMemoryAllocation;
            if (e != cudaSuccess) {
                rc = fail(cuda_code(e), "result buffers: %s", cudaGetErrorString(e));
                break;
            }
        }
        for (int i = 0; i < n && !rc; ++i) {
            cudaError_t e = cudaSetDevice(g_eng.devs[i]);
            if (e == cudaSuccess) e = bests[i].alloc((size_t)q_len * 4, g_eng.streams[i]);
            if (e == cudaSuccess && use_seed) e = deeps[i].alloc(q_len, g_eng.streams[i]);
            if (e == cudaSuccess && use_seed) e = cudaMemsetAsync(deeps[i].p, 0, q_len, g_eng.streams[i]);
            if (e != cudaSuccess) {
                rc = fail(cuda_code(e), "minima buffer: %s", cudaGetErrorString(e));
                break;
            }
            rc = k4b_best_init_device(bests[i].as<uint32_t>(), q_len, K, g_eng.streams[i]);
            if (rc) break;
            if (use_seed) {  // device i: its share of the index buckets, all probes
                k4b_seed_watch_depth(depth_cap, deeps[i].as<uint8_t>());
                rc = seed_run(qs[i], ts[i], both, clamp, core_len, 0, q_len, (uint32_t)i, (uint32_t)n,
                              bests[i].as<uint32_t>(), g_eng.streams[i], nullptr);
                k4b_seed_watch_depth(0, nullptr);
            } else
                rc = k4b_targeted_diag_device(qs[i], ts[i], both, clamp, (uint32_t)i, (uint32_t)n, bests[i].as<uint32_t>(),
                                              g_eng.streams[i], nullptr);
        }
        if (rc) break;
        trace.mark("alloc + enqueue");
        if (trace.on) {  // attribute the engine time (tracing only: adds a synchronisation)
            for (int i = 0; i < n; ++i) {
                cudaSetDevice(g_eng.devs[i]);
                cudaStreamSynchronize(g_eng.streams[i]);
            }
            trace.mark("engine kernels");
        }
        if (n > 1) {
            int nr = g_eng.nccl.GroupStart();
            for (int i = 0; i < n && !nr; ++i)
                nr = g_eng.nccl.AllReduce(bests[i].p, bests[i].p, q_len, kNcclUint32, kNcclMin, g_eng.comms[i], g_eng.streams[i]);
            int nr2 = g_eng.nccl.GroupEnd();
            if (!nr && !nr2 && use_seed) {  // the depth flags of the bucket shards: element-wise maximum
                nr = g_eng.nccl.GroupStart();
                for (int i = 0; i < n && !nr; ++i)
                    nr = g_eng.nccl.AllReduce(deeps[i].p, deeps[i].p, q_len, kNcclUint8, kNcclMax, g_eng.comms[i], g_eng.streams[i]);
                nr2 = g_eng.nccl.GroupEnd();
            }
            if (nr || nr2) {
                rc = fail(K4B_ERR_NCCL, "ncclAllReduce: %s", g_eng.nccl.GetErrorString(nr ? nr : nr2));
                break;
            }
        }
        cudaError_t e = cudaSetDevice(g_eng.devs[0]);
        unsigned long long *h_count = (unsigned long long *)((char *)h_out + (((size_t)q_len * 2 + 7) & ~(size_t)7));
        if (use_seed) {
            if (e == cudaSuccess) e = d_deep_count.alloc(8, g_eng.streams[0]);
            if (e == cudaSuccess) e = cudaMemsetAsync(d_deep_count.p, 0, 8, g_eng.streams[0]);
            if (e == cudaSuccess)
                e = launch_seed_deep_count(deeps[0].as<uint8_t>(), bests[0].as<uint32_t>(), qs[0]->view(), q_len, clamp,
                                           d_deep_count.as<unsigned long long>(), g_eng.streams[0]);
            if (e == cudaSuccess) e = cudaMemcpyAsync(h_count, d_deep_count.p, 8, cudaMemcpyDeviceToHost, g_eng.streams[0]);
        }
        if ((rc = k4b_targeted_finalize_device(qs[0], bests[0].as<uint32_t>(), clamp, d_out.as<uint16_t>(), g_eng.streams[0]))) break;
        if (use_seed)  // the wildcard probe K-mers the seed engine skipped
            for (size_t k = 0; k < impure.size() && !rc; ++k)
                rc = k4b_allpairs_min_device(qs[0], ts[0], both, same ? 1 : 0, impure[k].first, impure[k].second, clamp,
                                             d_out.as<uint16_t>() + impure[k].first, g_eng.streams[0], nullptr);
        if (rc) break;
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(h_out, d_out.p, (size_t)q_len * 2, cudaMemcpyDeviceToHost, g_eng.streams[0]);
        for (int i = 0; i < n; ++i) {
            cudaSetDevice(g_eng.devs[i]);
            const cudaError_t e2 = cudaStreamSynchronize(g_eng.streams[i]);
            if (e2 != cudaSuccess && e == cudaSuccess) e = e2;
        }
        if (e != cudaSuccess) {
            rc = fail(cuda_code(e), "targeted %s engine: %s", use_seed ? "seed" : "band", cudaGetErrorString(e));
            break;
        }
        trace.mark("kernels + D2H");
        for (uint32_t p = 0; p < q_len; ++p)
            if (h_out[p] <= K) out_h[p] = (uint8_t)h_out[p];
        if (use_seed) {
            g_depth.last_count = *h_count;
            g_depth.last_cap = depth_cap;
            g_depth.last_valid = true;
        }
    } while (0);
    for (int i = 0; i < n; ++i) {
        cudaSetDevice(g_eng.devs[i]);
        bests[i].release();
        deeps[i].release();
        if ((size_t)i < qs.size() && qs[i]) k4b_packed_free(qs[i]);
        if ((size_t)i < ts.size() && ts[i] && !((size_t)i < qs.size() && ts[i] == qs[i])) k4b_packed_free(ts[i]);
    }
    cudaSetDevice(g_eng.devs[0]);
    d_out.release();
    d_deep_count.release();
    trace.mark("free");
    return rc;
}

// auto: the band engine pays off once the pair matrix is large; tiny inputs stay on the
// all-pairs kernel (one launch)
static bool use_diag_engine(uint32_t len, uint32_t K) {
    const int e = engine_setting();
    if (e == 1) return false;
    if (e == 2) return len >= K;
    return len >= 200000u;
}

extern "C" int k4b_hamm_exhaustive_shard(const uint8_t *concat, uint32_t concat_len, uint32_t K,
                                         int both_strands, uint32_t q_begin, uint32_t q_end,
                                         uint16_t *out_min) {
    if (!concat || !out_min) return fail(K4B_ERR_PARAMS, "NULL buffer");
    RC(check_k(K, K4B_MIN_K, K4B_MAX_K));
    return run_sharded(concat, concat_len, nullptr, 0, K, both_strands, 1, q_begin, q_end, 0, 0,
                       SweepRange(), [&](uint32_t pos, uint16_t v) {
                           if (v <= K && v < out_min[pos]) out_min[pos] = v;
                       });
}

extern "C" int k4b_hamm_exhaustive(const uint8_t *concat, uint32_t concat_len, uint32_t K,
                                   int both_strands, uint32_t sweep_start, uint32_t sweep_end,
                                   uint16_t *out_min) {
    if (!concat || !out_min) return fail(K4B_ERR_PARAMS, "NULL buffer");
    RC(check_k(K, K4B_MIN_K, K4B_MAX_K));
    SweepRange sweep;
    RC(make_sweep(concat_len, K, sweep_start, sweep_end, sweep));
    if (!sweep.ranged && use_diag_engine(concat_len, K))
        return run_exhaustive_diag(concat, concat_len, K, both_strands, out_min);
    return run_sharded(concat, concat_len, nullptr, 0, K, both_strands, 1, 0, concat_len, 0, 1,
                       sweep, [&](uint32_t pos, uint16_t v) {
                           if (v <= K && v < out_min[pos]) out_min[pos] = v;
                       });
}

extern "C" int k4b_hamm_targeted_z(const uint8_t *target_concat, uint64_t target_len, uint32_t K, int R,
                                   int both_strands, int intra_inter_both, uint32_t q_begin,
                                   uint32_t q_end, uint8_t *out_h) {
    if (intra_inter_both < 0 || intra_inter_both > 2)
        return fail(K4B_ERR_PARAMS, "intra_inter_both=%d outside 0..2", intra_inter_both);
    if (!target_concat) return fail(K4B_ERR_PARAMS, "NULL buffer");
    RC(check_len(target_len));
    g_zf.mode = intra_inter_both;
    g_zf.starts.clear();
    if (intra_inter_both) {  // every entry is followed by EOS (SfxArray.cpp:1746-1750)
        g_zf.starts.push_back(0);
        for (uint64_t p = 0; p + 1 < target_len; ++p)
            if (target_concat[p] == 7) g_zf.starts.push_back((uint32_t)p + 1);
    }
    const int rc = k4b_hamm_targeted(target_concat, target_len, nullptr, 0, K, R, both_strands, q_begin, q_end, out_h);
    g_zf.mode = 0;
    g_zf.starts.clear();
    return rc;
}

extern "C" int k4b_hamm_targeted(const uint8_t *target_concat, uint64_t target_len,
                                 const uint8_t *probe_concat, uint32_t probe_len, uint32_t K,
                                 int R, int both_strands, uint32_t q_begin, uint32_t q_end,
                                 uint8_t *out_h) {
    if (!target_concat || !out_h) return fail(K4B_ERR_PARAMS, "NULL buffer");
    RC(check_k(K, K4B_MIN_K, 500));  // SfxArray.cpp:4255
    if (R < 1 || R > 10) return fail(K4B_ERR_PARAMS, "R=%d outside 1..10", R);
    if (K / (uint32_t)(R + 1) < 4)  // hammings.cpp:399-404
        return fail(K4B_ERR_PARAMS, "K/(R+1) must be >= 4");
    RC(check_len(target_len));
    // "not found" value of the pigeonhole search: CoreLen = K/(Rmax+1), K/CoreLen
    // (SfxArray.cpp:4462-4463)
    const uint32_t core = K / (uint32_t)(R + 1);
    const uint32_t notfound = K / core;
    if (!probe_concat) {
        // probes are the K-mers of the indexed assembly itself (CSfxArray::LocateSfxHammings,
        // SfxArray.cpp:4107-4220): exact self hits are skipped on the sense strand, the result
        // is additionally capped at 20 (:4208-4209)
        const uint32_t tl = (uint32_t)target_len;
        if (q_end == 0 || q_end > tl) q_end = tl;
        const int eng = engine_setting();
        if (q_begin == 0 && q_end == tl && (eng == 0 || eng == 3) && tl >= 4096) {
            int used = 0;
            const int rc = run_targeted_big(target_concat, tl, nullptr, 0, K, R, both_strands, std::min(notfound, 20u), core,
                                            true, false, out_h, &used);
            if (used || rc) return rc;
        }
        return run_sharded(target_concat, tl, nullptr, 0, K, both_strands, 1, q_begin, q_end,
                           std::min(notfound, 20u), 1, SweepRange(), [&](uint32_t pos, uint16_t v) {
                               if (v <= K) out_h[pos] = (uint8_t)v;
                           });
    }
    if (q_end == 0 || q_end > probe_len) q_end = probe_len;
    {
        const int eng = engine_setting();
        const bool whole = q_begin == 0 && q_end == probe_len;
        const bool big = (uint64_t)probe_len * target_len >= (1ull << 32);
        if (whole && eng != 1) {
            int used = 0;
            const int rc = run_targeted_big(target_concat, (uint32_t)target_len, probe_concat, probe_len, K, R, both_strands,
                                            notfound, core, eng == 0 || eng == 3, eng == 2 || (eng != 1 && big), out_h,
                                            &used);
            if (used || rc) return rc;
        }
    }
    return run_sharded(probe_concat, probe_len, target_concat, (uint32_t)target_len, K,
                       both_strands, 0, q_begin, q_end, notfound, 1, SweepRange(),
                       [&](uint32_t pos, uint16_t v) {
                           if (v <= K) out_h[pos] = (uint8_t)v;
                       });
}

// ------------------------------------------------------------------------------------------
// distribution of the minima
// ------------------------------------------------------------------------------------------
extern "C" int k4b_histogram_device(const uint16_t *d_min, uint32_t n, uint32_t K, unsigned long long *d_hist,
                                    void *stream) {
    if (!d_min || !d_hist) return fail(K4B_ERR_PARAMS, "NULL argument");
    RC(check_k(K, 1, 65533));
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaMemsetAsync(d_hist, 0, ((size_t)K + 2) * 8, st));
    CU(launch_histogram_u16(d_min, n, K + 2, d_hist, st));
    return K4B_OK;
}

extern "C" int k4b_hamm_histogram(const uint16_t *min, uint32_t n, uint32_t K, uint64_t *hist) {
    if (!min || !hist) return fail(K4B_ERR_PARAMS, "NULL buffer");
    RC(ensure_init());
    RC(check_k(K, 1, 65533));
    CU(cudaSetDevice(g_eng.devs[0]));
    cudaStream_t st = g_eng.streams[0];
    DevScratch d_v, d_h;
    CU(d_v.alloc((size_t)n * 2, st));
    CU(d_h.alloc(((size_t)K + 2) * 8, st));
    CU(cudaMemcpyAsync(d_v.p, min, (size_t)n * 2, cudaMemcpyHostToDevice, st));
    RC(k4b_histogram_device(d_v.as<uint16_t>(), n, K, d_h.as<unsigned long long>(), st));
    CU(cudaMemcpyAsync(hist, d_h.p, ((size_t)K + 2) * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return K4B_OK;
}

// ------------------------------------------------------------------------------------------
// integer-pipe microbenchmark
// ------------------------------------------------------------------------------------------
extern "C" int k4b_microbench_intpipe(int which, int iters, double *gops) {
    if (!gops) return fail(K4B_ERR_PARAMS, "gops is NULL");
    RC(ensure_init());
    struct Events {  // destroyed on every path
        cudaEvent_t a = nullptr, b = nullptr;
        ~Events() {
            if (a) cudaEventDestroy(a);
            if (b) cudaEventDestroy(b);
        }
    } ev;
    DevScratch sink;
    CU(sink.alloc(4, 0));
    CU(cudaEventCreate(&ev.a));
    CU(cudaEventCreate(&ev.b));
    int blocks = 0, threads = 0, ops = 0;
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {  // first rep is warm-up
        CU(cudaEventRecord(ev.a, 0));
        cudaError_t e = launch_microbench(which, iters, sink.as<uint32_t>(), &blocks, &threads, &ops, 0);
        if (e != cudaSuccess) return fail(K4B_ERR_CUDA, "microbench: %s", cudaGetErrorString(e));
        CU(cudaEventRecord(ev.b, 0));
        CU(cudaEventSynchronize(ev.b));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, ev.a, ev.b));
        const double g = (double)blocks * threads * ops * (double)iters / (ms * 1e-3) / 1e9;
        if (rep > 0 && g > best) best = g;
    }
    *gops = best;
    return K4B_OK;
}
