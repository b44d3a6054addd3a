// store.cu -- the C ABI of libshodh_gpu: device-resident vector store + exact search.
//
// Host-side mirror of the reference's VamanaIndex surface as RetrievalEngine uses it
// (reference src/vector_db/vamana.rs: search :764, add_vector :853, mark_deleted :813,
// len :976, build :200, extract_all_vectors; src/memory/retrieval.rs:53,912-918).
// Data layout in HBM: one dense row-major matrix [n_rows, dim] in the store dtype
// (fp16 by default: 768 B per 384-d row), 256-byte aligned, plus a soft-delete bitmap of
// one bit per row.  Ids are row numbers.  There is deliberately no CPU fallback.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <mutex>
#include <vector>

#include "sg_store_internal.h"

namespace sg {

// ---------------------------------------------------------------------------------------
// error string + launch counter
// ---------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---------------------------------------------------------------------------------------
// driver entry point for cuTensorMapEncodeTiled (no -lcuda link dependency)
// ---------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode_fn()
{
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    });
    return fn;
}

int encode_corpus_tmap(CUtensorMap *out, const void *d_rows, bool f16, uint32_t dim, uint64_t n_rows,
                       uint32_t box_rows)
{
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
        return SG_ERR_NO_DEVICE;
    }
    const uint32_t esize = f16 ? 2 : 4;
    cuuint64_t gdim[2] = {dim, n_rows ? n_rows : 1};
    cuuint64_t gstride[1] = {(cuuint64_t)dim * esize};
    cuuint32_t box[2] = {128u / esize, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                     const_cast<void *>(d_rows), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (dim=%u rows=%llu)", (int)r, dim,
                  (unsigned long long)n_rows);
        return SG_ERR_CUDA;
    }
    return SG_OK;
}

static int env_u32(const char *name, int dflt)
{
    const char *v = getenv(name);
    if (!v || !*v) return dflt;
    return atoi(v);
}

}  // namespace sg

using namespace sg;

namespace sg {

int check_device(int device, int *sm_count)
{
    // cudaGetDeviceProperties costs milliseconds: query each device once
    static std::mutex mu;
    static int cached_sms[64];
    static int cached_rc[64];
    static bool cached[64];
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error("no CUDA device available (%s); libshodh_gpu has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return SG_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n || device >= 64) {
        set_error("device ordinal %d out of range (0..%d)", device, n - 1);
        return SG_ERR_INVALID;
    }
    std::lock_guard<std::mutex> g(mu);
    if (!cached[device]) {
        cudaDeviceProp p;
        SG_CUDA_TRY(cudaGetDeviceProperties(&p, device));
        cached_sms[device] = p.multiProcessorCount;
        cached_rc[device] = p.major == 10 ? SG_OK : SG_ERR_NO_DEVICE;
        cached[device] = true;
        if (p.major != 10)
            set_error("device %d is sm_%d%d; libshodh_gpu is built for sm_100a (B200) only", device, p.major, p.minor);
    }
    if (cached_rc[device] != SG_OK) {
        set_error("device %d is not an sm_100 device; libshodh_gpu is built for sm_100a (B200) only", device);
        return cached_rc[device];
    }
    if (sm_count) *sm_count = cached_sms[device];
    return SG_OK;
}

int grow_pinned(sg_store *s, size_t bytes)
{
    if (bytes <= s->pinned_cap) return SG_OK;
    if (s->h_pinned) cudaFreeHost(s->h_pinned);
    s->h_pinned = nullptr;
    s->pinned_cap = 0;
    size_t ncap = bytes + bytes / 2;
    SG_CUDA_TRY(cudaHostAlloc(&s->h_pinned, ncap, cudaHostAllocDefault));
    s->pinned_cap = ncap;
    return SG_OK;
}

// ---- CUDA virtual memory management through the runtime's driver entry points (no -lcuda) ----
namespace {
struct VmmApi {
    CUresult (*reserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*addr_free)(CUdeviceptr, size_t) = nullptr;
    CUresult (*create)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
    CUresult (*release)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*unmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*set_access)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
    CUresult (*granularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags) = nullptr;
    bool ok = false;
};
const VmmApi &vmm_api()
{
    static VmmApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        auto get = [](const char *name, void **fn) {
            cudaDriverEntryPointQueryResult q;
            return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && *fn;
        };
        api.ok = get("cuMemAddressReserve", reinterpret_cast<void **>(&api.reserve)) &&
                 get("cuMemAddressFree", reinterpret_cast<void **>(&api.addr_free)) &&
                 get("cuMemCreate", reinterpret_cast<void **>(&api.create)) &&
                 get("cuMemRelease", reinterpret_cast<void **>(&api.release)) &&
                 get("cuMemMap", reinterpret_cast<void **>(&api.map)) &&
                 get("cuMemUnmap", reinterpret_cast<void **>(&api.unmap)) &&
                 get("cuMemSetAccess", reinterpret_cast<void **>(&api.set_access)) &&
                 get("cuMemGetAllocationGranularity", reinterpret_cast<void **>(&api.granularity));
        const char *v = getenv("SG_STORE_VMM");
        if (v && *v && atoi(v) == 0) api.ok = false;
    });
    return api;
}
CUmemAllocationProp vmm_prop(int device)
{
    CUmemAllocationProp p = {};
    p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    p.location.id = device;
    return p;
}
}  // namespace

void vmm_release(sg_store *s)
{
    if (!s->vmm) return;
    const VmmApi &api = vmm_api();
    if (s->mapped_bytes) api.unmap((CUdeviceptr)s->d_rows, s->mapped_bytes);
    for (unsigned long long h : s->vmm_handles) api.release((CUmemGenericAllocationHandle)h);
    if (s->d_rows) api.addr_free((CUdeviceptr)s->d_rows, s->va_bytes);
    s->vmm_handles.clear();
    s->vmm_sizes.clear();
    s->d_rows = nullptr;
    s->mapped_bytes = s->va_bytes = 0;
    s->vmm = false;
}

// Row storage that grows IN PLACE: a virtual range for the largest store this device could hold is reserved once,
// physical chunks are mapped at its end as rows arrive.  A live 10M-row store that grows needs no second buffer and
// no device-to-device copy, searches in flight keep reading valid memory, and the base address (tensor maps, graph
// kernels) never changes.  Returns false (caller falls back to allocate-and-copy) if VMM is unavailable.
static bool vmm_ensure(sg_store *s, uint64_t need_bytes, int *rc_out)
{
    const VmmApi &api = vmm_api();
    *rc_out = SG_OK;
    if (!api.ok || (s->d_rows && !s->vmm)) return false;
    const CUmemAllocationProp prop = vmm_prop(s->device);
    size_t gran = 0;
    if (api.granularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || gran == 0) return false;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return false;
    const uint64_t row_bytes = (uint64_t)s->dim * s->esize;
    uint64_t va_max = 0xFFFFFFFEull * row_bytes;                       // ids are u32
    if (va_max > (uint64_t)total_b) va_max = total_b;
    va_max = (va_max + gran - 1) / gran * gran;
    // The reservation is address space, not memory -- but address space is finite too (one index per user: a process may
    // hold thousands of stores), so a store starts with 16 GB (22M fp16 rows of 384) and, should it outgrow that, moves
    // its chunks into a range 8x as large: the physical chunks are re-mapped, nothing is copied.
    auto reserve = [&](uint64_t want_va, CUdeviceptr *base) -> uint64_t {
        uint64_t va = want_va > va_max ? va_max : want_va;
        va = (va + gran - 1) / gran * gran;
        return api.reserve(base, va, 0, 0, 0) == CUDA_SUCCESS ? va : 0;
    };
    const uint64_t va0 = (uint64_t)env_u32("SG_STORE_VA_MB", 16384) << 20;
    if (!s->d_rows) {
        CUdeviceptr base = 0;
        const uint64_t va = reserve(need_bytes * 2 > va0 ? need_bytes * 2 : va0, &base);
        if (!va) return false;
        s->d_rows = reinterpret_cast<void *>(base);
        s->va_bytes = va;
        s->mapped_bytes = 0;
        s->vmm = true;
    }
    if (need_bytes <= s->mapped_bytes) return true;
    if (need_bytes > va_max) {
        set_error("store would exceed the device's memory (%llu bytes of rows)", (unsigned long long)need_bytes);
        *rc_out = SG_ERR_OOM;
        return true;
    }
    if (need_bytes > s->va_bytes) {
        // move to a larger range: every search in flight must be done with the old addresses first
        if (cudaDeviceSynchronize() != cudaSuccess) {
            *rc_out = SG_ERR_CUDA;
            return true;
        }
        CUdeviceptr nbase = 0;
        uint64_t nva = reserve(s->va_bytes * 8 > need_bytes * 2 ? s->va_bytes * 8 : need_bytes * 2, &nbase);
        if (!nva) nva = reserve(need_bytes + need_bytes / 8, &nbase);
        if (!nva) {
            set_error("no address range of %llu bytes left for the store", (unsigned long long)need_bytes);
            *rc_out = SG_ERR_OOM;
            return true;
        }
        CUmemAccessDesc acc = {};
        acc.location = prop.location;
        acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        bool ok = s->mapped_bytes == 0 || api.unmap((CUdeviceptr)s->d_rows, s->mapped_bytes) == CUDA_SUCCESS;
        uint64_t off = 0;
        for (size_t i = 0; ok && i < s->vmm_handles.size(); ++i) {
            ok = api.map(nbase + off, s->vmm_sizes[i], 0, (CUmemGenericAllocationHandle)s->vmm_handles[i], 0) == CUDA_SUCCESS;
            off += s->vmm_sizes[i];
        }
        if (ok && off) ok = api.set_access(nbase, off, &acc, 1) == CUDA_SUCCESS;
        if (!ok) {
            set_error("re-mapping the store's %llu bytes into a larger address range failed", (unsigned long long)s->mapped_bytes);
            *rc_out = SG_ERR_CUDA;                      // the store is unusable: its rows are unmapped
            return true;
        }
        api.addr_free((CUdeviceptr)s->d_rows, s->va_bytes);
        s->d_rows = reinterpret_cast<void *>(nbase);
        s->va_bytes = nva;
    }
    // grow by at least 1/8 of what is mapped (amortises the map calls), in whole granules
    uint64_t want = need_bytes - s->mapped_bytes;
    if (want < s->mapped_bytes / 8) want = s->mapped_bytes / 8;
    want = (want + gran - 1) / gran * gran;
    if (s->mapped_bytes + want > s->va_bytes) want = s->va_bytes - s->mapped_bytes;
    CUmemGenericAllocationHandle h = 0;
    if (api.create(&h, want, &prop, 0) != CUDA_SUCCESS) {
        want = (need_bytes - s->mapped_bytes + gran - 1) / gran * gran;   // retry with the bare minimum
        if (api.create(&h, want, &prop, 0) != CUDA_SUCCESS) {
            set_error("out of device memory growing the store by %llu bytes", (unsigned long long)want);
            *rc_out = SG_ERR_OOM;
            return true;
        }
    }
    const CUdeviceptr at = (CUdeviceptr)s->d_rows + s->mapped_bytes;
    CUmemAccessDesc acc = {};
    acc.location = prop.location;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    if (api.map(at, want, 0, h, 0) != CUDA_SUCCESS || api.set_access(at, want, &acc, 1) != CUDA_SUCCESS) {
        api.release(h);
        set_error("mapping %llu bytes of row storage failed", (unsigned long long)want);
        *rc_out = SG_ERR_CUDA;
        return true;
    }
    s->vmm_handles.push_back((unsigned long long)h);
    s->vmm_sizes.push_back(want);
    s->mapped_bytes += want;
    return true;
}

int ensure_rows(sg_store *s, uint64_t need_rows)
{
    if (need_rows <= s->cap) return SG_OK;
    if (need_rows >= 0xFFFFFFFFull) {
        set_error("store is limited to 2^32-2 rows (ids are u32)");
        return SG_ERR_UNSUPPORTED;
    }
    const size_t row_bytes = (size_t)s->dim * s->esize;
    // the soft-delete bitmap is small (1 bit per row): reallocated + copied, generously sized
    auto grow_bitmap = [&](uint64_t ncap) -> int {
        const size_t words = ncap / 32 + 1;
        if (words <= s->h_deleted.size() && s->d_deleted) return SG_OK;
        uint32_t *nbm = nullptr;
        SG_CUDA_TRY(cudaMalloc(&nbm, words * 4));
        cudaError_t e = cudaMemsetAsync(nbm, 0, words * 4, s->stream);
        if (e == cudaSuccess && s->n && s->d_deleted)
            e = cudaMemcpyAsync(nbm, s->d_deleted, (s->n / 32 + 1) * 4, cudaMemcpyDeviceToDevice, s->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
        if (e != cudaSuccess) {
            cudaFree(nbm);
            set_error("growing the deleted-row bitmap failed: %s", cudaGetErrorString(e));
            return SG_ERR_CUDA;
        }
        if (s->d_deleted) cudaFree(s->d_deleted);   // synchronises the device: no search in flight still reads it
        s->d_deleted = nbm;
        s->h_deleted.resize(words, 0u);
        return SG_OK;
    };
    {   // preferred: grow in place
        int vrc = SG_OK;
        if (vmm_ensure(s, need_rows * row_bytes, &vrc)) {
            if (vrc != SG_OK) return vrc;
            const uint64_t ncap = s->mapped_bytes / row_bytes;
            int rc = grow_bitmap(ncap + ncap / 2);
            if (rc != SG_OK) return rc;
            s->cap = ncap;
            return SG_OK;
        }
    }
    // fallback: allocate x1.5, copy, free
    uint64_t ncap = s->cap + s->cap / 2;
    if (ncap < need_rows) ncap = need_rows;
    ncap = (ncap + 1023) & ~1023ull;
    void *nrows = nullptr;
    SG_CUDA_TRY(cudaMalloc(&nrows, ncap * row_bytes));
    cudaError_t e = cudaSuccess;
    if (s->n) e = cudaMemcpyAsync(nrows, s->d_rows, s->n * row_bytes, cudaMemcpyDeviceToDevice, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    if (e != cudaSuccess) {
        cudaFree(nrows);
        set_error("growing the store failed: %s", cudaGetErrorString(e));
        return SG_ERR_CUDA;
    }
    int rc = grow_bitmap(ncap);
    if (rc != SG_OK) {
        cudaFree(nrows);
        return rc;
    }
    if (s->d_rows) cudaFree(s->d_rows);
    s->d_rows = nrows;
    s->cap = ncap;
    return SG_OK;
}

}  // namespace sg

namespace {

struct ScanPlan {
    uint32_t W, ring, cap, grid;
    size_t smem[5];  // indexed by qb (1,2,4)
};

// choose warps-per-CTA / ring depth / buffer capacity so the CTA fits in 227 KB
int plan_scan(const sg_store *s, uint32_t k, uint32_t qb, ScanPlan *p)
{
    const uint32_t nbox = s->dim * s->esize / 128;
    uint32_t kp = 32;
    while (kp < k) kp <<= 1;
    p->cap = 2 * kp;
    // ring depth per warp: 6 boxes (one 384-d fp16 row tile) for multi-query sweeps; 4 measured ~5 % faster
    // for the single-query sweep at 10M rows (1.10 / 1.12 ms vs 1.16 / 1.18 ms for k = 10 / 100)
    uint32_t ring_dflt = nbox < 6 ? (nbox < 3 ? 3 : nbox) : 6;
    if (qb == 1 && ring_dflt > 4) ring_dflt = 4;
    uint32_t ring = (uint32_t)env_u32("SG_SCAN_RING", (int)ring_dflt);
    if (ring < 2) ring = 2;
    uint32_t Wmax = (uint32_t)env_u32("SG_SCAN_WARPS", 8);
    if (Wmax > 16) Wmax = 16;
    const size_t limit = 227 * 1024;
    for (uint32_t W = Wmax; W >= 1; --W) {
        size_t sm = scan_smem_bytes(W, ring, qb, s->dim, p->cap);
        uint32_t M = 2;
        while (M < W * k) M <<= 1;
        if (sm <= limit && (size_t)M * 8 <= (size_t)W * ring * 4096) {
            p->W = W;
            p->ring = ring;
            p->smem[qb] = sm;
            const uint64_t n_tiles = (s->n + 31) / 32;
            uint64_t g = (n_tiles + W - 1) / W;
            p->grid = (uint32_t)(g < (uint64_t)s->sm_count ? (g ? g : 1) : s->sm_count);
            return SG_OK;
        }
    }
    set_error("no scan configuration fits shared memory (dim=%u k=%u qb=%u)", s->dim, k, qb);
    return SG_ERR_UNSUPPORTED;
}

}  // namespace

namespace sg {

// exact path A.  all pointers are device pointers; runs asynchronously on st
int search_exact_locked(sg_store *s, const float *d_q, uint32_t b, uint32_t k, uint32_t *d_ids, float *d_dist,
                        uint32_t *d_counts, cudaStream_t st, bool record, bool use_deleted, const PushArgs *push)
{
    if (s->n == 0) {  // VamanaIndex::search on an empty index: Ok(vec![]) (vamana.rs:766-768)
        SG_CUDA_TRY(launch_merge_keys(nullptr, b, 0, k, s->id_base, nullptr, d_ids, d_dist, d_counts, st, push));
        return SG_OK;
    }
    // one plan for the widest sweep; narrower sweeps reuse W/ring/grid (their smem is smaller)
    // NormalizedDotProduct sweeps exist for 1, 2 and 4 queries, the other metrics for 1 and 4
    const bool nd = s->metric == SG_METRIC_NEG_DOT;
    const uint32_t qb_max = b >= (nd ? 3u : 2u) ? 4 : b;
    ScanPlan plan;
    int rc = plan_scan(s, k, qb_max, &plan);
    if (rc != SG_OK) return rc;
    const uint32_t L = plan.grid;
    rc = grow(s->d_lists, s->lists_cap, (size_t)b * L * k);
    if (rc != SG_OK) return rc;
    const uint32_t groups = merge_groups(L, k);
    if (groups > 1 && (rc = grow(s->d_lists2, s->lists2_cap, (size_t)b * groups * k)) != SG_OK) return rc;

    CUtensorMap tmap;
    rc = encode_corpus_tmap(&tmap, s->d_rows, s->dtype == SG_DTYPE_F16, s->dim, s->n, 32);
    if (rc != SG_OK) return rc;

    const bool prof = record && s->profiling && 2 * (size_t)(s->ev_count + 1) <= s->ev.size();
    if (prof) SG_CUDA_TRY(cudaEventRecord(s->ev[2 * s->ev_count], st));
    uint32_t done = 0;
    while (done < b) {
        const uint32_t rem = b - done;
        const uint32_t qb = rem >= (nd ? 3u : 2u) ? 4 : rem;
        const uint32_t nq = rem < qb ? rem : qb;
        ScanArgs a;
        a.queries = d_q + (size_t)done * s->dim;
        a.nq = nq;
        a.dim = s->dim;
        a.n_rows = s->n;
        a.deleted = (s->n_deleted && use_deleted) ? s->d_deleted : nullptr;
        a.k = k;
        a.cap = plan.cap;
        a.nbox = s->dim * s->esize / 128;
        a.ring = plan.ring;
        a.out_lists = s->d_lists + (size_t)done * L * k;
        a.flag_count = nullptr;
        a.flag_list = nullptr;
        size_t smem = scan_smem_bytes(plan.W, plan.ring, qb, s->dim, plan.cap);
        cudaError_t e = launch_scan_exact(tmap, a, (int)qb, s->dtype == SG_DTYPE_F16, s->order, s->metric, L, plan.W, smem, st);
        if (e != cudaSuccess) {
            set_error("scan kernel launch failed: %s", cudaGetErrorString(e));
            return SG_ERR_CUDA;
        }
        ++s->last_scan_launches;
        done += nq;
    }
    if (prof) {
        SG_CUDA_TRY(cudaEventRecord(s->ev[2 * s->ev_count + 1], st));
        ++s->ev_count;
    }
    SG_CUDA_TRY(launch_merge_keys(s->d_lists, b, L, k, s->id_base, groups > 1 ? s->d_lists2 : nullptr, d_ids, d_dist,
                                  d_counts, st, push));
    return SG_OK;
}

int store_truncate_locked(sg_store *s)
{
    SG_CUDA_TRY(cudaSetDevice(s->device));
    SG_CUDA_TRY(cudaStreamSynchronize(s->stream));
    s->n = 0;
    s->n_deleted = 0;
    s->graph_len = 0;
    std::fill(s->h_deleted.begin(), s->h_deleted.end(), 0u);
    if (s->d_deleted) SG_CUDA_TRY(cudaMemsetAsync(s->d_deleted, 0, s->h_deleted.size() * 4, s->stream));
    SG_CUDA_TRY(cudaMemsetAsync(s->d_max_norm, 0, sizeof(float), s->stream));
    SG_CUDA_TRY(cudaStreamSynchronize(s->stream));
    return SG_OK;
}

// see sg_store_internal.h
int chain_begin(sg_store *s, cudaStream_t st)
{
    if (s->chain_pending && s->chain_stream != st) SG_CUDA_TRY(cudaStreamWaitEvent(st, s->ev_chain, 0));
    return SG_OK;
}
int chain_end(sg_store *s, cudaStream_t st)
{
    if (!s->ev_chain) SG_CUDA_TRY(cudaEventCreateWithFlags(&s->ev_chain, cudaEventDisableTiming));
    SG_CUDA_TRY(cudaEventRecord(s->ev_chain, st));
    s->chain_stream = st;
    s->chain_pending = true;
    return SG_OK;
}

}  // namespace sg

namespace {


// tensor-core path B: approximate sweep + exact re-score + certification; uncertified queries are
// re-run on path A.
struct TcPlan {
    uint32_t kk, stages, kbs, units, n_lists, grid, J, cap_l, qtile;
    bool pair;
};

// candidates re-scored per query = k + slack; fewer => fewer appends and a cheaper finalize, more slack =>
// fewer exact re-runs on tie-heavy corpora.  The reference's own call shape is k = 4 * limit with
// limit = 3 * max_results (retrieval.rs:912-918, memory/mod.rs:3585): k = 120 for a top-10 recall.
uint32_t tc_kk(uint32_t k)
{
    if (k <= 10) return 16;
    if (k <= 24) return 32;
    return (k + 32 + 7) / 8 * 8;
}

bool tc_plan(const sg_store *s, uint32_t b, uint32_t k, TcPlan *p)
{
    if (s->dtype != SG_DTYPE_F16 || s->dim % 64 != 0 || s->metric != SG_METRIC_NEG_DOT) return false;
    if (s->tc_mode == 1 || env_u32("SG_TC_DISABLE", 0)) return false;
    if (b < (uint32_t)env_u32("SG_TC_MIN_B", 3)) return false;    // b = 3, 4: one tensor-core sweep (1.24 ms at 10M) beats the 4-query exact sweep (1.79 ms)
    if (k > 128 || s->n < 4096) return false;
    p->kk = tc_kk(k);
    const uint32_t kk_env = (uint32_t)env_u32("SG_TC_KK", 0);
    if (kk_env >= k + 2 && kk_env <= TC_KK_MAX) p->kk = kk_env;
    // CTA-pair form (256 queries per sweep, tensor-bound: ~1.56 ms per 10M rows whatever b is) from SG_TC_PAIR_MIN_B
    // queries; below, the 1-CTA form (128 per sweep).  Measured at 10M rows, b = 128 / 96 / 64: 1.33 / 1.44 / 1.48 ms with
    // the 1-CTA form against 1.44-1.51 / 1.61 / 1.35 ms with the pair (run-to-run spread ~10 %): no clear winner below
    // 129, so the pair starts where it is the only form that serves the batch in one sweep.
    const int pair_env = env_u32("SG_TC_PAIR", 1);
    p->pair = pair_env == 2 || (pair_env == 1 && b >= (uint32_t)env_u32("SG_TC_PAIR_MIN_B", 129));
    p->qtile = p->pair ? 256 : 128;
    const uint32_t rows_tile = p->pair ? 256 : 128;
    const uint64_t n_tiles = (s->n + rows_tile - 1) / rows_tile;
    const uint32_t units_max = p->pair ? (uint32_t)s->sm_count / 2 : (uint32_t)s->sm_count;   // clusters or CTAs
    p->units = (uint32_t)(n_tiles < units_max ? n_tiles : units_max);
    p->grid = p->pair ? 2 * p->units : p->units;
    // every CTA (pair) keeps TWO candidate regions per query (one per epilogue warp group, each scanning every other
    // 32-row chunk).  The sample: a unit contributes the J <= 4 best chunk maxima of its first tile.
    p->n_lists = 2 * p->units;
    p->J = (p->kk + p->units - 1) / p->units;
    if (p->J > 4 || p->J > rows_tile / 32) return false;
    p->cap_l = p->kk <= 32 ? 48 : 128;
    const uint32_t cap_env = (uint32_t)env_u32("SG_TC_CAP", 0);
    if (cap_env >= 8 && cap_env <= 4096) p->cap_l = cap_env;
    // ring: stages of kbs K-blocks (2 when the row has an even number of 64-element K-blocks), as many as fit
    p->kbs = ((s->dim / 64) % 2 == 0 && env_u32("SG_TC_KBS", 2) >= 2) ? 2 : 1;
    const uint32_t st_max = (uint32_t)env_u32("SG_TC_STAGES", 8) / p->kbs;
    for (uint32_t stages = st_max < 2 ? 2 : (st_max > 8 ? 8 : st_max); stages >= 2; --stages) {
        if (stages * p->kbs < 3) break;
        if (tc_smem_bytes(s->dim, stages, p->kbs) <= 227 * 1024) {
            p->stages = stages;
            return true;
        }
    }
    return false;
}

constexpr uint32_t TC_B_CHUNK = 1024;    // queries per pipeline pass (bounds the candidate scratch)

int search_tc_chunk(sg_store *s, const TcPlan &pl, const float *d_q, uint32_t b, uint32_t k, uint32_t *d_ids,
                    float *d_dist, uint32_t *d_counts, cudaStream_t st, const PushArgs *push)
{
    const uint32_t qtile = pl.qtile, n_lists = pl.n_lists, kk = pl.kk;
    const uint32_t b_pad = (b + qtile - 1) / qtile * qtile;
    const uint32_t n_qtiles = b_pad / qtile;
    int rc;
    if (b_pad > s->qmeta_cap) {
        size_t c[4] = {s->qmeta_cap, s->qmeta_cap, s->qmeta_cap, s->qmeta_cap * TC_HIST};
        if ((rc = grow(s->d_flag_list, c[0], b_pad)) != SG_OK) return rc;
        if ((rc = grow(s->d_gtau, c[1], b_pad)) != SG_OK) return rc;
        if ((rc = grow(s->d_flags, c[2], b_pad)) != SG_OK) return rc;
        if ((rc = grow(s->d_hist, c[3], (size_t)c[0] * TC_HIST)) != SG_OK) return rc;
        s->qmeta_cap = c[0];
    }
    if (!s->d_arrive) {                              // monotonic arrival counter of the sweeps' meeting point
        SG_CUDA_TRY(cudaMalloc(&s->d_arrive, 4));
        SG_CUDA_TRY(cudaMemset(s->d_arrive, 0, 4));
        s->arrive_total = 0;
    }
    if ((rc = grow(s->d_cand, s->cand_cap_q, (size_t)b_pad * n_lists * pl.cap_l)) != SG_OK) return rc;
    if ((rc = grow(s->d_cnts, s->cnts_cap, (size_t)b_pad * n_lists)) != SG_OK) return rc;
    const uint32_t nv_pad = (pl.units * pl.J + 3) / 4 * 4;
    if ((rc = grow(s->d_maxima, s->maxima_cap, (size_t)b_pad * nv_pad)) != SG_OK) return rc;
    CUtensorMap tmap_rows;
    if ((rc = encode_corpus_tmap(&tmap_rows, s->d_rows, true, s->dim, s->n, 128)) != SG_OK) return rc;

    const size_t smem = tc_smem_bytes(s->dim, pl.stages, pl.kbs);
    unsigned long long *d_trace = nullptr;           // SG_TC_TRACE=1: phase stamps of the sweep, printed to stderr (debug)
    const bool trace = env_u32("SG_TC_TRACE", 0) != 0;
    if (trace) {
        SG_CUDA_TRY(cudaMalloc(&d_trace, (size_t)pl.grid * 16 * 8));
        SG_CUDA_TRY(cudaMemsetAsync(d_trace, 0, (size_t)pl.grid * 16 * 8, st));
    }
    const bool prof = s->profiling && 2 * (size_t)(s->ev_count + 1) <= s->ev.size();
    if (prof) SG_CUDA_TRY(cudaEventRecord(s->ev[2 * s->ev_count], st));
    for (uint32_t qt = 0; qt < n_qtiles; ++qt) {
        TcArgs a{};
        a.n_rows = s->n;
        a.dim = s->dim;
        a.nq = b - qt * qtile < qtile ? b - qt * qtile : qtile;
        a.q_tile = qt;
        a.stages = pl.stages;
        a.kk = kk;
        a.b_pad = b_pad;
        a.J = pl.J;
        a.cap_l = pl.cap_l;
        a.kbs = pl.kbs;
        a.deleted = s->n_deleted ? s->d_deleted : nullptr;
        a.gtau = s->d_gtau;
        a.maxima = s->d_maxima;
        a.nv_pad = nv_pad;
        a.arrive = s->d_arrive;
        s->arrive_total += pl.grid;
        a.arrive_target = s->arrive_total;
        a.queries = d_q;
        a.qtile = qtile;
        a.nflag = s->d_nflag;
        a.nv = pl.units * pl.J;
        a.hist = s->d_hist;
        a.cand = s->d_cand;
        a.cnts = s->d_cnts;
        a.trace = d_trace;
        cudaError_t e = pl.pair ? launch_scan_tc2(tmap_rows, a, pl.grid, smem, st) : launch_scan_tc(tmap_rows, a, pl.grid, smem, st);
        if (e != cudaSuccess) {
            set_error("tensor-core scan launch failed: %s", cudaGetErrorString(e));
            return SG_ERR_CUDA;
        }
        ++s->last_scan_launches;
    }
    if (prof) {
        SG_CUDA_TRY(cudaEventRecord(s->ev[2 * s->ev_count + 1], st));
        ++s->ev_count;
    }
    if (trace) {
        std::vector<unsigned long long> h((size_t)pl.grid * 16);
        SG_CUDA_TRY(cudaStreamSynchronize(st));
        SG_CUDA_TRY(cudaMemcpy(h.data(), d_trace, h.size() * 8, cudaMemcpyDeviceToHost));
        cudaFree(d_trace);
        unsigned long long t0 = ~0ull, mx[12] = {0}, mn_done = ~0ull;
        for (uint32_t c = 0; c < pl.grid; ++c)
            if (h[(size_t)c * 16] && h[(size_t)c * 16] < t0) t0 = h[(size_t)c * 16];
        for (uint32_t c = 0; c < pl.grid; ++c) {
            for (int i = 0; i < 12; ++i)
                if (h[(size_t)c * 16 + i] > mx[i]) mx[i] = h[(size_t)c * 16 + i];
            if (h[(size_t)c * 16 + 11] && h[(size_t)c * 16 + 11] < mn_done) mn_done = h[(size_t)c * 16 + 11];
        }
        auto us = [&](unsigned long long t) { return t >= t0 ? (t - t0) / 1e3 : -1.0; };
        fprintf(stderr, "[sg tc trace] us since the first CTA's epilogue start (max over CTAs): enter %.1f | sample acc ready %.1f | "
                        "sample scanned %.1f | all arrived %.1f | seeds %.1f | tile 1 %.1f | 2 %.1f | 4 %.1f | 8 %.1f | 16 %.1f | "
                        "32 %.1f | done %.1f (first CTA done %.1f)\n",
                us(mx[0]), us(mx[1]), us(mx[2]), us(mx[3]), us(mx[4]), us(mx[5]), us(mx[6]), us(mx[7]), us(mx[8]), us(mx[9]),
                us(mx[10]), us(mx[11]), us(mn_done));
    }
    FinalizeArgs f{};
    f.queries = d_q;
    f.rows = s->d_rows;
    f.dim = s->dim;
    f.k = k;
    f.kk = kk;
    f.order = (uint32_t)s->order;
    f.id_base = s->id_base;
    f.n_live = s->n - s->n_deleted;
    f.cand = s->d_cand;
    f.cnts = s->d_cnts;
    f.n_lists = n_lists;
    f.cap_l = pl.cap_l;
    f.gtau = s->d_gtau;
    f.max_row_norm = s->d_max_norm;
    f.c_acc = tc_c_acc(s->dim);
    {   // rows staged per re-score round: q + rows in dynamic shared memory (18.5 KB more are static).  Up to 32 rows within
        // 28 KB for the small over-selections (k <= 24: one round, 4 CTAs per SM); up to 64 rows within 52 KB when kk is
        // large (k = 120 -> kk = 152: 3 rounds instead of 5 -- a round is one global round trip + one sequential 384-step
        // chain per candidate, the rounds do not overlap)
        const size_t row_b = (size_t)(s->dim / 2 + 1) * 4, q_b = (size_t)s->dim * 4;
        const size_t budget = kk > 32 ? 52 * 1024 : 28 * 1024, cap = kk > 32 ? 64 : 32;
        size_t rcn = (budget - q_b) / row_b;
        f.rc = (uint32_t)(rcn > cap ? cap : (rcn < 1 ? 1 : rcn));
    }
    f.ids_out = d_ids;
    f.dist_out = d_dist;
    f.counts_out = d_counts;
    f.flags = s->d_flags;
    f.n_flagged = s->d_nflag;
    f.flag_list = s->d_flag_list;
    f.flag_total = s->d_flag_total;
    if (push) f.push = *push;
    SG_CUDA_TRY(launch_finalize_tc(f, b, st));

    // Certification outcome, handled ON THE DEVICE (no host round trip, the call stays asynchronous):
    // the exact sweep is launched in "flagged" mode -- it reads the number and indices of uncertified
    // queries from device memory, exits at once when there are none (the normal case, ~3 us), and
    // otherwise re-runs them 4 per sweep; the flagged merge then overwrites exactly those result rows
    // (and, multi-GPU, pushes them to the peers: finalize only pushed the certified ones).
    s->tc_searches += b;
    ScanPlan plan;
    if ((rc = plan_scan(s, k, 4, &plan)) != SG_OK) return rc;
    const uint32_t LA = plan.grid;
    if ((rc = grow(s->d_lists, s->lists_cap, (size_t)b * LA * k)) != SG_OK) return rc;
    CUtensorMap tmap32;
    if ((rc = encode_corpus_tmap(&tmap32, s->d_rows, true, s->dim, s->n, 32)) != SG_OK) return rc;
    ScanArgs sa;
    sa.queries = d_q;
    sa.nq = 0;
    sa.dim = s->dim;
    sa.n_rows = s->n;
    sa.deleted = s->n_deleted ? s->d_deleted : nullptr;
    sa.k = k;
    sa.cap = plan.cap;
    sa.nbox = s->dim * s->esize / 128;
    sa.ring = plan.ring;
    sa.out_lists = s->d_lists;
    sa.flag_count = s->d_nflag;
    sa.flag_list = s->d_flag_list;
    cudaError_t e2 = launch_scan_exact(tmap32, sa, 4, true, s->order, SG_METRIC_NEG_DOT, LA, plan.W,
                                       scan_smem_bytes(plan.W, plan.ring, 4, s->dim, plan.cap), st);
    if (e2 != cudaSuccess) {
        set_error("flagged exact sweep launch failed: %s", cudaGetErrorString(e2));
        return SG_ERR_CUDA;
    }
    SG_CUDA_TRY(launch_merge_keys_flagged(s->d_lists, b, LA, k, s->id_base, s->d_nflag, s->d_flag_list, d_ids, d_dist,
                                          d_counts, st, push));
    return SG_OK;
}

// all pointers are device pointers.  push (multi-GPU) = hand every query's final local result to the peers
int search_device_locked(sg_store *s, const float *d_q, uint32_t b, uint32_t k, uint32_t *d_ids, float *d_dist,
                         uint32_t *d_counts, cudaStream_t st, const PushArgs *push = nullptr)
{
    if (b == 0) return SG_OK;
    if (k == 0 || k > 1024) {
        set_error("k=%u outside the supported range 1..1024", k);
        return k == 0 ? SG_ERR_INVALID : SG_ERR_UNSUPPORTED;
    }
    s->last_scan_launches = 0;
    TcPlan pl;
    if (!tc_plan(s, b, k, &pl)) return search_exact_locked(s, d_q, b, k, d_ids, d_dist, d_counts, st, true, true, push);
    for (uint32_t off = 0; off < b; off += TC_B_CHUNK) {
        const uint32_t bb = b - off < TC_B_CHUNK ? b - off : TC_B_CHUNK;
        PushArgs pc{};
        if (push) {
            pc = *push;
            pc.q0 = push->q0 + off;
        }
        int rc = search_tc_chunk(s, pl, d_q + (size_t)off * s->dim, bb, k, d_ids + (size_t)off * k, d_dist + (size_t)off * k,
                                 d_counts + off, st, push ? &pc : nullptr);
        if (rc != SG_OK) return rc;
    }
    return SG_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
extern "C" {

const char *sg_version(void) { return "shodh_gpu 0.1.0 (sm_100a)"; }
const char *sg_last_error(void) { return g_err; }
uint64_t sg_kernel_launch_count(void) { return g_launches.load(); }

int sg_device_count(int *count_out)
{
    if (!count_out) return SG_ERR_INVALID;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count_out = 0;
        set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
        return SG_ERR_NO_DEVICE;
    }
    *count_out = n;
    return SG_OK;
}

int sg_store_create(sg_store **out, int device, uint32_t dim, int dtype, int metric)
{
    if (!out) return SG_ERR_INVALID;
    *out = nullptr;
    if (dtype != SG_DTYPE_F16 && dtype != SG_DTYPE_F32) {
        set_error("unknown dtype %d", dtype);
        return SG_ERR_INVALID;
    }
    if (metric < SG_METRIC_NEG_DOT || metric > SG_METRIC_COSINE_SIM) {
        set_error("unknown metric %d", metric);
        return SG_ERR_INVALID;
    }
    const uint32_t mult = dtype == SG_DTYPE_F16 ? 64 : 32;
    if (dim == 0 || dim % mult != 0 || dim > 2048) {
        set_error("dim=%u unsupported: must be a non-zero multiple of %u and <= 2048", dim, mult);
        return SG_ERR_UNSUPPORTED;
    }
    int sms = 0;
    int rc = check_device(device, &sms);
    if (rc != SG_OK) return rc;
    SG_CUDA_TRY(cudaSetDevice(device));
    sg_store *s = new (std::nothrow) sg_store();
    if (!s) return SG_ERR_OOM;
    s->device = device;
    s->sm_count = sms;
    s->dim = dim;
    s->dtype = dtype;
    s->esize = dtype == SG_DTYPE_F16 ? 2 : 4;
    s->metric = metric;
    cudaError_t e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_max_norm, sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&s->d_nflag, sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&s->d_flag_total, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(s->d_flag_total, 0, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(s->d_max_norm, 0, sizeof(float));
    if (e != cudaSuccess) {
        set_error("stream/event creation failed: %s", cudaGetErrorString(e));
        delete s;
        return SG_ERR_CUDA;
    }
    *out = s;
    return SG_OK;
}

void sg_store_destroy(sg_store *s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->vmm) vmm_release(s);
    else cudaFree(s->d_rows);
    cudaFree(s->d_deleted);
    cudaFree(s->d_queries);
    cudaFree(s->d_lists);
    cudaFree(s->d_lists2);
    cudaFree(s->d_resblk);
    cudaFree(s->d_ids);
    cudaFree(s->d_dist);
    cudaFree(s->d_counts);
    cudaFree(s->d_stage);
    cudaFree(s->d_vec2mem);
    cudaFree(s->d_mem_out);
    cudaFree(s->d_sim_out);
    cudaFree(s->d_memcnt);
    cudaFree(s->d_max_norm);
    cudaFree(s->d_maxima);
    cudaFree(s->d_gtau);
    cudaFree(s->d_flags);
    cudaFree(s->d_nflag);
    cudaFree(s->d_flag_list);
    cudaFree(s->d_flag_total);
    cudaFree(s->d_cand);
    cudaFree(s->d_hist);
    cudaFree(s->d_arrive);
    cudaFree(s->d_cnts);
    if (s->ev_chain) cudaEventDestroy(s->ev_chain);
    graph_free(s);
    if (s->h_pinned) cudaFreeHost(s->h_pinned);
    for (cudaEvent_t e : s->ev) cudaEventDestroy(e);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

int sg_store_reserve(sg_store *s, uint64_t n_rows)
{
    if (!s) return SG_ERR_INVALID;
    std::lock_guard<std::mutex> g(s->mu);
    SG_CUDA_TRY(cudaSetDevice(s->device));
    { int rc_ = chain_begin(s, s->stream); if (rc_ != SG_OK) return rc_; }
    return ensure_rows(s, n_rows);
}

int sg_store_append(sg_store *s, const float *rows, uint64_t n, uint32_t *first_id_out)
{
    if (!s || (!rows && n)) return SG_ERR_INVALID;
    std::lock_guard<std::mutex> g(s->mu);
    SG_CUDA_TRY(cudaSetDevice(s->device));
    { int rc_ = chain_begin(s, s->stream); if (rc_ != SG_OK) return rc_; }
    if (first_id_out) *first_id_out = (uint32_t)s->n;
    if (n == 0) return SG_OK;
    int rc = ensure_rows(s, s->n + n);
    if (rc != SG_OK) return rc;
    const size_t chunk_rows = (size_t)(16u << 20) / ((size_t)s->dim * 4) ? (size_t)(16u << 20) / ((size_t)s->dim * 4) : 1;
    rc = grow_pinned(s, chunk_rows * s->dim * 4);
    if (rc != SG_OK) return rc;
    if (s->dtype == SG_DTYPE_F16) {
        rc = grow(s->d_stage, s->stage_cap, chunk_rows * s->dim);
        if (rc != SG_OK) return rc;
    }
    uint64_t done = 0;
    while (done < n) {
        const uint64_t cnt = (n - done) < chunk_rows ? (n - done) : chunk_rows;
        const size_t elems = (size_t)cnt * s->dim;
        memcpy(s->h_pinned, rows + done * s->dim, elems * 4);
        uint8_t *dst = static_cast<uint8_t *>(s->d_rows) + (s->n + done) * (size_t)s->dim * s->esize;
        if (s->dtype == SG_DTYPE_F16) {
            SG_CUDA_TRY(cudaMemcpyAsync(s->d_stage, s->h_pinned, elems * 4, cudaMemcpyHostToDevice, s->stream));
            SG_CUDA_TRY(launch_f32_to_f16(s->d_stage, reinterpret_cast<__half *>(dst), elems, s->stream));
        } else {
            SG_CUDA_TRY(cudaMemcpyAsync(dst, s->h_pinned, elems * 4, cudaMemcpyHostToDevice, s->stream));
        }
        SG_CUDA_TRY(launch_max_row_norm(dst, s->dtype == SG_DTYPE_F16, cnt, s->dim, s->d_max_norm, s->stream));
        SG_CUDA_TRY(cudaStreamSynchronize(s->stream));
        done += cnt;
    }
    const uint64_t first = s->n;
    s->n += n;
    return graph_append_locked(s, first, n);   // VamanaIndex::add_vector links each new row when a graph exists
}

int sg_store_append_device(sg_store *s, const float *d_rows, uint64_t n, uint32_t *first_id_out, void *cuda_stream)
{
    if (!s || (!d_rows && n)) return SG_ERR_INVALID;
    std::lock_guard<std::mutex> g(s->mu);
    SG_CUDA_TRY(cudaSetDevice(s->device));
    if (first_id_out) *first_id_out = (uint32_t)s->n;
    if (n == 0) return SG_OK;
    int rc = ensure_rows(s, s->n + n);
    if (rc != SG_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    if ((rc = chain_begin(s, st)) != SG_OK) return rc;
    const size_t elems = (size_t)n * s->dim;
    uint8_t *dst = static_cast<uint8_t *>(s->d_rows) + s->n * (size_t)s->dim * s->esize;
    if (s->dtype == SG_DTYPE_F16) {
        SG_CUDA_TRY(launch_f32_to_f16(d_rows, reinterpret_cast<__half *>(dst), elems, st));
    } else {
        SG_CUDA_TRY(cudaMemcpyAsync(dst, d_rows, elems * 4, cudaMemcpyDeviceToDevice, st));
    }
    SG_CUDA_TRY(launch_max_row_norm(dst, s->dtype == SG_DTYPE_F16, n, s->dim, s->d_max_norm, st));
    SG_CUDA_TRY(cudaStreamSynchronize(st));
    const uint64_t first = s->n;
    s->n += n;
    return graph_append_locked(s, first, n);
}

int sg_store_append_synthetic(sg_store *s, uint64_t seed, uint64_t first_row, uint64_t n)
{
    if (!s) return SG_ERR_INVALID;
    std::lock_guard<std::mutex> g(s->mu);
    SG_CUDA_TRY(cudaSetDevice(s->device));
    { int rc_ = chain_begin(s, s->stream); if (rc_ != SG_OK) return rc_; }
    int rc = ensure_rows(s, s->n + n);
    if (rc != SG_OK) return rc;
    uint8_t *dst = static_cast<uint8_t *>(s->d_rows) + s->n * (size_t)s->dim * s->esize;
    SG_CUDA_TRY(launch_synth_rows(dst, s->dtype == SG_DTYPE_F16, seed, first_row, n, s->dim, s->stream));
    SG_CUDA_TRY(launch_max_row_norm(dst, s->dtype == SG_DTYPE_F16, n, s->dim, s->d_max_norm, s->stream));
    SG_CUDA_TRY(cudaStreamSynchronize(s->stream));
    const uint64_t first = s->n;
    s->n += n;
    return graph_append_locked(s, first, n);
}

int sg_store_mark_deleted(sg_store *s, uint32_t id, int *ok_out)
{
    if (!s) return SG_ERR_INVALID;
    std::lock_guard<std::mutex> g(s->mu);
    if (id >= s->n) {  // mark_deleted returns false for unknown ids (vamana.rs:814-819)
        if (ok_out) *ok_out = 0;
        return SG_OK;
    }
    SG_CUDA_TRY(cudaSetDevice(s->device));
    { int rc_ = chain_begin(s, s->stream); if (rc_ != SG_OK) return rc_; }
    uint32_t &w = s->h_deleted[id >> 5];
    const uint32_t bit = 1u << (id & 31);
    if (!(w & bit)) {
        w |= bit;
        ++s->n_deleted;
        SG_CUDA_TRY(cudaMemcpyAsync(s->d_deleted + (id >> 5), &w, 4, cudaMemcpyHostToDevice, s->stream));
        SG_CUDA_TRY(cudaStreamSynchronize(s->stream));
    }
    if (ok_out) *ok_out = 1;
    return SG_OK;
}

int sg_store_is_deleted(sg_store *s, uint32_t id, int *deleted_out)
{
    if (!s || !deleted_out) return SG_ERR_INVALID;
    std::lock_guard<std::mutex> g(s->mu);
    *deleted_out = (id < s->n) ? (int)((s->h_deleted[id >> 5] >> (id & 31)) & 1u) : 0;
    return SG_OK;
}

int sg_store_clear_deleted(sg_store *s)
{
    if (!s) return SG_ERR_INVALID;
    std::lock_guard<std::mutex> g(s->mu);
    if (!s->d_deleted) return SG_OK;
    SG_CUDA_TRY(cudaSetDevice(s->device));
    { int rc_ = chain_begin(s, s->stream); if (rc_ != SG_OK) return rc_; }
    std::fill(s->h_deleted.begin(), s->h_deleted.end(), 0u);
    s->n_deleted = 0;
    SG_CUDA_TRY(cudaMemsetAsync(s->d_deleted, 0, s->h_deleted.size() * 4, s->stream));
    SG_CUDA_TRY(cudaStreamSynchronize(s->stream));
    return SG_OK;
}

int sg_store_memory_info(sg_store *s, int *in_place_out, uint64_t *capacity_rows_out, uint64_t *bytes_out,
                         uint32_t *chunks_out)
{
    if (!s) return SG_ERR_INVALID;
    std::lock_guard<std::mutex> g(s->mu);
    if (in_place_out) *in_place_out = s->vmm ? 1 : 0;
    if (capacity_rows_out) *capacity_rows_out = s->cap;
    if (bytes_out) *bytes_out = s->vmm ? s->mapped_bytes : s->cap * (uint64_t)s->dim * s->esize;
    if (chunks_out) *chunks_out = s->vmm ? (uint32_t)s->vmm_handles.size() : (s->d_rows ? 1u : 0u);
    return SG_OK;
}

int sg_store_len(const sg_store *s, uint64_t *n_out)
{
    if (!s || !n_out) return SG_ERR_INVALID;
    std::lock_guard<std::mutex> g(s->mu);
    *n_out = s->n;
    return SG_OK;
}

int sg_store_dim(const sg_store *s, uint32_t *dim_out)
{
    if (!s || !dim_out) return SG_ERR_INVALID;
    *dim_out = s->dim;
    return SG_OK;
}

int sg_store_metric(const sg_store *s, int *metric_out)
{
    if (!s || !metric_out) return SG_ERR_INVALID;
    *metric_out = s->metric;
    return SG_OK;
}

int sg_store_deleted_count(const sg_store *s, uint64_t *n_out)
{
    if (!s || !n_out) return SG_ERR_INVALID;
    std::lock_guard<std::mutex> g(s->mu);
    *n_out = s->n_deleted;
    return SG_OK;
}

int sg_store_get_vectors(sg_store *s, uint64_t first, uint64_t n, float *out)
{
    if (!s || (!out && n)) return SG_ERR_INVALID;
    std::lock_guard<std::mutex> g(s->mu);
    if (first + n > s->n) {
        set_error("rows [%llu,%llu) out of range (len %llu)", (unsigned long long)first,
                  (unsigned long long)(first + n), (unsigned long long)s->n);
        return SG_ERR_INVALID;
    }
    if (n == 0) return SG_OK;
    SG_CUDA_TRY(cudaSetDevice(s->device));
    { int rc_ = chain_begin(s, s->stream); if (rc_ != SG_OK) return rc_; }
    const size_t chunk_rows = (size_t)(16u << 20) / ((size_t)s->dim * 4) ? (size_t)(16u << 20) / ((size_t)s->dim * 4) : 1;
    int rc = grow_pinned(s, chunk_rows * s->dim * 4);
    if (rc != SG_OK) return rc;
    if (s->dtype == SG_DTYPE_F16) {
        rc = grow(s->d_stage, s->stage_cap, chunk_rows * s->dim);
        if (rc != SG_OK) return rc;
    }
    uint64_t done = 0;
    while (done < n) {
        const uint64_t cnt = (n - done) < chunk_rows ? (n - done) : chunk_rows;
        const size_t elems = (size_t)cnt * s->dim;
        const uint8_t *src = static_cast<const uint8_t *>(s->d_rows) + (first + done) * (size_t)s->dim * s->esize;
        if (s->dtype == SG_DTYPE_F16) {
            SG_CUDA_TRY(launch_f16_to_f32(reinterpret_cast<const __half *>(src), s->d_stage, elems, s->stream));
            SG_CUDA_TRY(cudaMemcpyAsync(s->h_pinned, s->d_stage, elems * 4, cudaMemcpyDeviceToHost, s->stream));
        } else {
            SG_CUDA_TRY(cudaMemcpyAsync(s->h_pinned, src, elems * 4, cudaMemcpyDeviceToHost, s->stream));
        }
        SG_CUDA_TRY(cudaStreamSynchronize(s->stream));
        memcpy(out + done * s->dim, s->h_pinned, elems * 4);
        done += cnt;
    }
    return SG_OK;
}

int sg_store_set_order(sg_store *s, int order)
{
    if (!s || (order != SG_ORDER_SCALAR && order != SG_ORDER_AVX2)) return SG_ERR_INVALID;
    std::lock_guard<std::mutex> g(s->mu);
    s->order = order;
    return SG_OK;
}

int sg_store_set_id_base(sg_store *s, uint32_t id_base)
{
    if (!s) return SG_ERR_INVALID;
    std::lock_guard<std::mutex> g(s->mu);
    s->id_base = id_base;
    return SG_OK;
}

int sg_store_set_profiling(sg_store *s, int on)
{
    if (!s) return SG_ERR_INVALID;
    std::lock_guard<std::mutex> g(s->mu);
    SG_CUDA_TRY(cudaSetDevice(s->device));
    s->profiling = on != 0;
    s->ev_count = 0;
    const size_t want = 2 * 4096;
    while (s->profiling && s->ev.size() < want) {
        cudaEvent_t e;
        SG_CUDA_TRY(cudaEventCreate(&e));
        s->ev.push_back(e);
    }
    return SG_OK;
}

int sg_store_profile_read(sg_store *s, float *ms_out, uint32_t cap, uint32_t *n_out, uint32_t *launches_out)
{
    if (!s || !n_out || (cap && !ms_out)) return SG_ERR_INVALID;
    std::lock_guard<std::mutex> g(s->mu);
    SG_CUDA_TRY(cudaSetDevice(s->device));
    const uint32_t n = s->ev_count;
    *n_out = n;
    if (launches_out) *launches_out = s->last_scan_launches;
    if (n) SG_CUDA_TRY(cudaEventSynchronize(s->ev[2 * (n - 1) + 1]));
    for (uint32_t i = 0; i < n && i < cap; ++i)
        SG_CUDA_TRY(cudaEventElapsedTime(ms_out + i, s->ev[2 * i], s->ev[2 * i + 1]));
    s->ev_count = 0;
    return SG_OK;
}

int sg_store_last_scan_ms(sg_store *s, float *ms_out, uint32_t *launches_out)
{
    if (!s || !ms_out) return SG_ERR_INVALID;
    std::lock_guard<std::mutex> g(s->mu);
    if (!s->ev_count) {
        set_error("no profiled search recorded (sg_store_set_profiling(s,1), then search)");
        return SG_ERR_INVALID;
    }
    SG_CUDA_TRY(cudaSetDevice(s->device));
    const uint32_t i = s->ev_count - 1;
    SG_CUDA_TRY(cudaEventSynchronize(s->ev[2 * i + 1]));
    SG_CUDA_TRY(cudaEventElapsedTime(ms_out, s->ev[2 * i], s->ev[2 * i + 1]));
    if (launches_out) *launches_out = s->last_scan_launches;
    s->ev_count = 0;
    return SG_OK;
}

int sg_store_tc_stats(sg_store *s, uint64_t *queries_out, uint64_t *flagged_out)
{
    if (!s) return SG_ERR_INVALID;
    std::lock_guard<std::mutex> g(s->mu);
    if (queries_out) *queries_out = s->tc_searches;
    if (flagged_out) {
        SG_CUDA_TRY(cudaSetDevice(s->device));
        unsigned long long v = 0;
        SG_CUDA_TRY(cudaMemcpy(&v, s->d_flag_total, sizeof(v), cudaMemcpyDeviceToHost));   // synchronises the device
        *flagged_out = v;
    }
    return SG_OK;
}

int sg_store_search_device(sg_store *s, const float *d_q, uint32_t b, uint32_t k, uint32_t *d_ids, float *d_dist,
                           uint32_t *d_counts, void *cuda_stream)
{
    if (!s || (b && (!d_q || !d_ids || !d_dist || !d_counts))) return SG_ERR_INVALID;
    std::lock_guard<std::mutex> g(s->mu);
    SG_CUDA_TRY(cudaSetDevice(s->device));
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);  // NULL = the CUDA (legacy) default stream
    int rc = chain_begin(s, st);
    if (rc != SG_OK) return rc;
    rc = search_device_locked(s, d_q, b, k, d_ids, d_dist, d_counts, st);
    const int rc2 = chain_end(s, st);
    return rc != SG_OK ? rc : rc2;
}

int sg_store_search_sharded_device(sg_store *s, const float *d_q, uint32_t b, uint32_t k, const uint64_t *peer_buffers,
                                   uint32_t rank, uint32_t world, uint32_t epoch, uint32_t *d_ids_out, float *d_dist_out,
                                   uint32_t *d_counts_out, void *cuda_stream)
{
    if (!s || !peer_buffers || world == 0 || world > 8 || rank >= world || epoch == 0) {
        set_error("sg_store_search_sharded_device: bad arguments (world 1..8, rank < world, epoch >= 1)");
        return SG_ERR_INVALID;
    }
    if (b && (!d_q || !d_ids_out || !d_dist_out || !d_counts_out)) return SG_ERR_INVALID;
    if (b == 0) return SG_OK;
    if (k == 0 || k > 1024) {
        set_error("k=%u outside the supported range 1..1024", k);
        return k == 0 ? SG_ERR_INVALID : SG_ERR_UNSUPPORTED;
    }
    std::lock_guard<std::mutex> g(s->mu);
    SG_CUDA_TRY(cudaSetDevice(s->device));
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    int rc = chain_begin(s, st);
    if (rc != SG_OK) return rc;
    // this shard's own (ids, dist)[b,k]: the store's result scratch
    if ((size_t)b * k > s->res_cap) {
        size_t c1 = s->res_cap, c2 = s->res_cap;
        if ((rc = grow(s->d_ids, c1, (size_t)b * k)) != SG_OK) return rc;
        if ((rc = grow(s->d_dist, c2, (size_t)b * k)) != SG_OK) return rc;
        s->res_cap = c1 < c2 ? c1 : c2;
    }
    if ((rc = grow(s->d_counts, s->cnt_cap, (size_t)b)) != SG_OK) return rc;
    PushArgs p{};
    for (uint32_t r = 0; r < world; ++r) p.peer[r] = reinterpret_cast<uint32_t *>(peer_buffers[r]);
    p.rank = rank; p.world = world; p.epoch = epoch; p.b = b; p.k = k; p.q0 = 0;
    rc = search_device_locked(s, d_q, b, k, s->d_ids, s->d_dist, s->d_counts, st, &p);
    if (rc == SG_OK) rc = exchange_wait_merge(s->device, p, d_ids_out, d_dist_out, d_counts_out, st);
    const int rc2 = chain_end(s, st);
    return rc != SG_OK ? rc : rc2;
}

int sg_store_set_tc_mode(sg_store *s, int mode)
{
    if (!s || mode < 0 || mode > 1) return SG_ERR_INVALID;
    std::lock_guard<std::mutex> g(s->mu);
    s->tc_mode = mode;
    return SG_OK;
}

// ---- diagnostics (tests only): raw tensor-core scores and exact scores of a row range -----------------
int sg_debug_tc_scores(sg_store *s, const float *queries, uint32_t b, uint64_t first_row, uint64_t n_rows, float *out)
{
    if (!s || !queries || !out || b == 0 || b > 256 || n_rows == 0) return SG_ERR_INVALID;
    std::lock_guard<std::mutex> g(s->mu);
    SG_CUDA_TRY(cudaSetDevice(s->device));
    if (s->dtype != SG_DTYPE_F16 || s->dim % 64 != 0 || first_row + n_rows > s->n || tc_smem_bytes(s->dim, 3, 1) > 227 * 1024) {
        set_error("sg_debug_tc_scores: needs an f16 store, dim %% 64 == 0, the query block within shared memory (dim <= 704) and a "
                  "row range inside the store");
        return SG_ERR_UNSUPPORTED;
    }
    int rc = chain_begin(s, s->stream);
    if (rc != SG_OK) return rc;
    const bool pair = b > 128;
    const uint32_t qtile = pair ? 256 : 128;
    float *d_q = nullptr, *d_out = nullptr;
    cudaError_t e = cudaMalloc(&d_q, (size_t)b * s->dim * 4);
    if (e == cudaSuccess) e = cudaMalloc(&d_out, (size_t)b * n_rows * 4);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_q, queries, (size_t)b * s->dim * 4, cudaMemcpyHostToDevice, s->stream);
    rc = SG_OK;
    if (e == cudaSuccess) {
        const uint32_t rows_tile = pair ? 256 : 128;
        const uint64_t n_tiles = (s->n + rows_tile - 1) / rows_tile;
        const uint32_t units = pair ? (uint32_t)s->sm_count / 2 : (uint32_t)s->sm_count;
        const uint32_t n_lists = (uint32_t)(n_tiles < units ? n_tiles : units);
        uint32_t stages = 8;
        while (stages > 3 && tc_smem_bytes(s->dim, stages, 1) > 227 * 1024) --stages;
        CUtensorMap tmap_rows;
        rc = encode_corpus_tmap(&tmap_rows, s->d_rows, true, s->dim, s->n, 128);
        if (rc == SG_OK) {
            TcArgs a{};
            a.n_rows = s->n;
            a.dim = s->dim;
            a.nq = b;
            a.stages = stages;
            a.kbs = 1;
            a.queries = d_q;
            a.qtile = qtile;
            a.dump = d_out;
            a.dump_row0 = first_row;
            a.dump_rows = n_rows;
            e = pair ? launch_scan_tc2(tmap_rows, a, 2 * n_lists, tc_smem_bytes(s->dim, stages, 1), s->stream)
                     : launch_scan_tc(tmap_rows, a, n_lists, tc_smem_bytes(s->dim, stages, 1), s->stream);
        }
    }
    if (rc == SG_OK && e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, (size_t)b * n_rows * 4, cudaMemcpyDeviceToHost, s->stream);
    if (rc == SG_OK && e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cudaFree(d_q); cudaFree(d_out);
    if (rc == SG_OK && e != cudaSuccess) {
        set_error("sg_debug_tc_scores: %s", cudaGetErrorString(e));
        rc = SG_ERR_CUDA;
    }
    return rc;
}

int sg_debug_exact_scores(sg_store *s, const float *queries, uint32_t b, uint64_t first_row, uint64_t n_rows, float *out)
{
    if (!s || !queries || !out || b == 0 || n_rows == 0) return SG_ERR_INVALID;
    std::lock_guard<std::mutex> g(s->mu);
    SG_CUDA_TRY(cudaSetDevice(s->device));
    if (s->dtype != SG_DTYPE_F16 || first_row + n_rows > s->n) {
        set_error("sg_debug_exact_scores: needs an f16 store and a row range inside the store");
        return SG_ERR_UNSUPPORTED;
    }
    int rc = chain_begin(s, s->stream);
    if (rc != SG_OK) return rc;
    float *d_q = nullptr, *d_out = nullptr;
    cudaError_t e = cudaMalloc(&d_q, (size_t)b * s->dim * 4);
    if (e == cudaSuccess) e = cudaMalloc(&d_out, (size_t)b * n_rows * 4);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_q, queries, (size_t)b * s->dim * 4, cudaMemcpyHostToDevice, s->stream);
    if (e == cudaSuccess) e = launch_exact_scores(d_q, b, s->d_rows, s->dim, first_row, n_rows, s->order, d_out, s->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, (size_t)b * n_rows * 4, cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cudaFree(d_q); cudaFree(d_out);
    if (e != cudaSuccess) {
        set_error("sg_debug_exact_scores: %s", cudaGetErrorString(e));
        return SG_ERR_CUDA;
    }
    return SG_OK;
}

int sg_store_search(sg_store *s, const float *queries, uint32_t b, uint32_t k, uint32_t *ids_out, float *dist_out,
                    uint32_t *counts_out)
{
    if (!s || (b && (!queries || !ids_out || !dist_out || !counts_out))) return SG_ERR_INVALID;
    if (b == 0) return SG_OK;
    if (k == 0) {
        set_error("k must be >= 1");
        return SG_ERR_INVALID;
    }
    std::lock_guard<std::mutex> g(s->mu);
    // SG_HOST_TRACE=1: host-clock breakdown of this call, averaged over 64 calls, to stderr (debug)
    static const bool htrace = env_u32("SG_HOST_TRACE", 0) != 0;
    static double hacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    static int hn = 0;
    static cudaEvent_t hev[4] = {nullptr, nullptr, nullptr, nullptr};
    if (htrace && !hev[0])
        for (auto &e : hev) cudaEventCreate(&e);
    auto now = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double h0 = htrace ? now() : 0.0;
    SG_CUDA_TRY(cudaSetDevice(s->device));
    const size_t qbytes = (size_t)b * s->dim * 4, rbytes = (size_t)b * k * 4, cbytes = (size_t)b * 4;
    int rc = grow_pinned(s, qbytes + 2 * rbytes + cbytes);
    if (rc != SG_OK) return rc;
    if ((rc = grow(s->d_queries, s->q_cap, (size_t)b * s->dim)) != SG_OK) return rc;
    // ids | distances | counts in ONE device block -> one device-to-host copy per search
    const size_t rwords = (size_t)b * k;
    if ((rc = grow(s->d_resblk, s->resblk_cap, 2 * rwords + b)) != SG_OK) return rc;
    uint32_t *d_ids = s->d_resblk;
    float *d_dist = reinterpret_cast<float *>(s->d_resblk + rwords);
    uint32_t *d_counts = s->d_resblk + 2 * rwords;

    uint8_t *hp = static_cast<uint8_t *>(s->h_pinned);
    // a caller that keeps its queries in page-locked memory (cudaHostAlloc / cudaHostRegister / torch pin_memory) is
    // DMA'd from directly; pageable memory goes through the store's pinned staging buffer first
    const void *h2d_src = hp;
    {
        cudaPointerAttributes pa{};
        if (cudaPointerGetAttributes(&pa, queries) == cudaSuccess && pa.type == cudaMemoryTypeHost) h2d_src = queries;
        else {
            (void)cudaGetLastError();
            memcpy(hp, queries, qbytes);
        }
    }
    const double h1 = htrace ? now() : 0.0;
    if ((rc = chain_begin(s, s->stream)) != SG_OK) return rc;
    if (htrace) cudaEventRecord(hev[0], s->stream);
    SG_CUDA_TRY(cudaMemcpyAsync(s->d_queries, h2d_src, qbytes, cudaMemcpyHostToDevice, s->stream));
    if (htrace) cudaEventRecord(hev[1], s->stream);
    const double h2 = htrace ? now() : 0.0;
    rc = search_device_locked(s, s->d_queries, b, k, d_ids, d_dist, d_counts, s->stream);
    if (rc != SG_OK) return rc;
    if (htrace) cudaEventRecord(hev[2], s->stream);
    SG_CUDA_TRY(cudaMemcpyAsync(hp + qbytes, s->d_resblk, 2 * rbytes + cbytes, cudaMemcpyDeviceToHost, s->stream));
    if (htrace) cudaEventRecord(hev[3], s->stream);
    const double h3 = htrace ? now() : 0.0;
    SG_CUDA_TRY(cudaStreamSynchronize(s->stream));
    const double h4 = htrace ? now() : 0.0;
    memcpy(ids_out, hp + qbytes, rbytes);
    memcpy(dist_out, hp + qbytes + rbytes, rbytes);
    memcpy(counts_out, hp + qbytes + 2 * rbytes, cbytes);
    if (htrace) {
        const double h5 = now();
        hacc[0] += h1 - h0; hacc[1] += h2 - h1; hacc[2] += h3 - h2; hacc[3] += h4 - h3; hacc[4] += h5 - h4;
        float g01 = 0, g12 = 0, g23 = 0;
        cudaEventElapsedTime(&g01, hev[0], hev[1]); cudaEventElapsedTime(&g12, hev[1], hev[2]); cudaEventElapsedTime(&g23, hev[2], hev[3]);
        hacc[5] += g01 * 1e3; hacc[6] += g12 * 1e3; hacc[7] += g23 * 1e3;
        if (++hn == 64) {
            fprintf(stderr, "[sg host trace] sg_store_search b=%u k=%u, us per call: setup+staging %.1f | H2D enqueue %.1f | kernel launches + D2H enqueue %.1f | "
                            "wait for the GPU %.1f | copy-out %.1f || GPU timeline: H2D %.1f | kernels %.1f | D2H %.1f\n", b, k, hacc[0] / hn,
                    hacc[1] / hn, hacc[2] / hn, hacc[3] / hn, hacc[4] / hn, hacc[5] / hn, hacc[6] / hn, hacc[7] / hn);
            hn = 0;
            for (double &x : hacc) x = 0;
        }
    }
    return SG_OK;
}

int sg_top_k_similar(int device, const float *query, uint32_t query_dim, const float *candidates, uint64_t n,
                     uint32_t dim, uint32_t k, int order, uint32_t *pos_out, float *score_out, uint32_t *n_out)
{
    if (!n_out || (n && k && (!query || !candidates || !pos_out || !score_out))) return SG_ERR_INVALID;
    *n_out = 0;
    if (n == 0 || k == 0) return SG_OK;
    const uint32_t r = (uint32_t)(n < k ? n : k);
    if (query_dim != dim) {
        // cosine_similarity returns 0.0 on a length mismatch (similarity.rs:11-13): every score ties at 0 and
        // the stable sort keeps the input order
        for (uint32_t i = 0; i < r; ++i) {
            pos_out[i] = i;
            score_out[i] = 0.0f;
        }
        *n_out = r;
        return SG_OK;
    }
    if (k > 1024) {
        set_error("sg_top_k_similar: k=%u outside the supported range 1..1024", k);
        return SG_ERR_UNSUPPORTED;
    }
    sg_store *st = nullptr;
    int rc = sg_store_create(&st, device, dim, SG_DTYPE_F32, SG_METRIC_COSINE_SIM);
    if (rc != SG_OK) return rc;
    if ((rc = sg_store_set_order(st, order)) == SG_OK && (rc = sg_store_append(st, candidates, n, nullptr)) == SG_OK) {
        uint32_t cnt = 0;
        rc = sg_store_search(st, query, 1, k, pos_out, score_out, &cnt);
        if (rc == SG_OK) {
            for (uint32_t i = 0; i < cnt; ++i) score_out[i] = -score_out[i];   // the store ranks by -similarity
            *n_out = cnt;
        }
    }
    sg_store_destroy(st);
    return rc;
}

int sg_store_set_memory_ids(sg_store *s, const uint64_t *vec2mem, uint64_t n)
{
    if (!s || (n && !vec2mem)) return SG_ERR_INVALID;
    std::lock_guard<std::mutex> g(s->mu);
    SG_CUDA_TRY(cudaSetDevice(s->device));
    int rc = grow(s->d_vec2mem, s->map_cap, (size_t)(n ? n : 1) * 2);
    if (rc != SG_OK) return rc;
    if (n) SG_CUDA_TRY(cudaMemcpy(s->d_vec2mem, vec2mem, (size_t)n * 16, cudaMemcpyHostToDevice));
    s->n_map = n;
    return SG_OK;
}

int sg_store_search_memory_ids(sg_store *s, const float *queries, uint32_t b, uint32_t limit, uint64_t *mem_out,
                               float *sim_out, uint32_t *counts_out)
{
    if (!s || (b && (!queries || !mem_out || !sim_out || !counts_out))) return SG_ERR_INVALID;
    if (b == 0) return SG_OK;
    // retrieval.rs:915-916: k = limit * VECTOR_SEARCH_CANDIDATE_MULTIPLIER (2, constants.rs:350) * 2
    const uint64_t k64 = (uint64_t)limit * 4;
    if (limit == 0 || k64 > 1024) {
        set_error("limit=%u: the scan is asked for 4*limit candidates, supported up to 1024", limit);
        return limit == 0 ? SG_ERR_INVALID : SG_ERR_UNSUPPORTED;
    }
    const uint32_t k = (uint32_t)k64;
    std::lock_guard<std::mutex> g(s->mu);
    SG_CUDA_TRY(cudaSetDevice(s->device));
    const size_t qbytes = (size_t)b * s->dim * 4;
    int rc = grow_pinned(s, qbytes + (size_t)b * limit * 20 + (size_t)b * 4);
    if (rc != SG_OK) return rc;
    if ((rc = grow(s->d_queries, s->q_cap, (size_t)b * s->dim)) != SG_OK) return rc;
    if ((size_t)b * k > s->res_cap) {
        size_t c1 = s->res_cap, c2 = s->res_cap;
        if ((rc = grow(s->d_ids, c1, (size_t)b * k)) != SG_OK) return rc;
        if ((rc = grow(s->d_dist, c2, (size_t)b * k)) != SG_OK) return rc;
        s->res_cap = c1 < c2 ? c1 : c2;
    }
    if ((rc = grow(s->d_counts, s->cnt_cap, (size_t)b)) != SG_OK) return rc;
    if ((size_t)b * limit > s->post_cap) {
        size_t c1 = s->post_cap, c2 = s->post_cap;
        uint64_t *m = s->d_mem_out;
        size_t cm = c1 * 2;
        if ((rc = grow(m, cm, (size_t)b * limit * 2)) != SG_OK) return rc;
        s->d_mem_out = m;
        if ((rc = grow(s->d_sim_out, c2, (size_t)b * limit)) != SG_OK) return rc;
        s->post_cap = c2 < cm / 2 ? c2 : cm / 2;
    }
    if ((rc = grow(s->d_memcnt, s->postc_cap, (size_t)b)) != SG_OK) return rc;
    uint8_t *hp = static_cast<uint8_t *>(s->h_pinned);
    memcpy(hp, queries, qbytes);
    if ((rc = chain_begin(s, s->stream)) != SG_OK) return rc;
    SG_CUDA_TRY(cudaMemcpyAsync(s->d_queries, hp, qbytes, cudaMemcpyHostToDevice, s->stream));
    rc = search_device_locked(s, s->d_queries, b, k, s->d_ids, s->d_dist, s->d_counts, s->stream);
    if (rc != SG_OK) return rc;
    SG_CUDA_TRY(launch_search_ids_post(s->d_ids, s->d_dist, s->d_counts, b, k, s->d_vec2mem, s->n_map, s->id_base, limit,
                                       s->d_mem_out, s->d_sim_out, s->d_memcnt, s->stream));
    const size_t mb = (size_t)b * limit * 16, sb = (size_t)b * limit * 4;
    SG_CUDA_TRY(cudaMemcpyAsync(hp + qbytes, s->d_mem_out, mb, cudaMemcpyDeviceToHost, s->stream));
    SG_CUDA_TRY(cudaMemcpyAsync(hp + qbytes + mb, s->d_sim_out, sb, cudaMemcpyDeviceToHost, s->stream));
    SG_CUDA_TRY(cudaMemcpyAsync(hp + qbytes + mb + sb, s->d_memcnt, (size_t)b * 4, cudaMemcpyDeviceToHost, s->stream));
    SG_CUDA_TRY(cudaStreamSynchronize(s->stream));
    memcpy(mem_out, hp + qbytes, mb);
    memcpy(sim_out, hp + qbytes + mb, sb);
    memcpy(counts_out, hp + qbytes + mb + sb, (size_t)b * 4);
    return SG_OK;
}

int sg_merge_topk_device_strided(int device, const uint32_t *d_ids, const float *d_dist, uint32_t b, uint32_t n_lists,
                                 uint32_t k, uint64_t list_stride, uint64_t query_stride, uint32_t *d_ids_out,
                                 float *d_dist_out, uint32_t *d_counts_out, void *cuda_stream)
{
    if (b && (!d_ids || !d_dist || !d_ids_out || !d_dist_out || !d_counts_out)) return SG_ERR_INVALID;
    if (k == 0 || k > 1024) {
        set_error("k=%u outside the supported range 1..1024", k);
        return SG_ERR_UNSUPPORTED;
    }
    int rc = check_device(device, nullptr);
    if (rc != SG_OK) return rc;
    SG_CUDA_TRY(cudaSetDevice(device));
    SG_CUDA_TRY(launch_merge_pairs(d_ids, d_dist, b, n_lists, k, list_stride, query_stride, d_ids_out, d_dist_out,
                                   d_counts_out, static_cast<cudaStream_t>(cuda_stream)));
    return SG_OK;
}

int sg_merge_topk_device(int device, const uint32_t *d_ids, const float *d_dist, uint32_t b, uint32_t n_lists,
                         uint32_t k, uint32_t *d_ids_out, float *d_dist_out, uint32_t *d_counts_out, void *cuda_stream)
{
    return sg_merge_topk_device_strided(device, d_ids, d_dist, b, n_lists, k, k, (uint64_t)n_lists * k, d_ids_out,
                                        d_dist_out, d_counts_out, cuda_stream);
}

// ---------------------------------------------------------------------------------------
// relevance fusion
// ---------------------------------------------------------------------------------------
void sg_fuse_default_weights(sg_fuse_weights *w)
{
    if (!w) return;  // relevance.rs:64-88
    w->semantic = 0.18f;
    w->entity = 0.17f;
    w->tag = 0.05f;
    w->importance = 0.05f;
    w->momentum = 0.28f;
    w->access_count = 0.14f;
    w->graph_strength = 0.13f;
}

void sg_fuse_default_config(sg_fuse_config *c)
{
    if (!c) return;  // relevance.rs:147-177, :914
    c->recency_boost_hours = 24;
    c->recency_boost_multiplier = 1.2f;
    c->graph_boost_multiplier = 1.15f;
    c->min_score = 0.25f;
    c->max_results = 5;
}

}  // extern "C"

namespace {
// Per-device transient arena for the small host-buffer calls (sg_fuse_topk, sg_rrf_fuse): they sit on the
// batch-1 latency path, where a cudaMalloc + cudaHostAlloc + stream create per call would cost more than the
// kernel.  Grown on demand, kept for the life of the process; calls on one device serialise on the mutex.
struct CallScratch {
    std::mutex mu;
    uint8_t *d = nullptr, *h = nullptr;
    size_t cap = 0;
    cudaStream_t st = nullptr;
};
CallScratch g_call_scratch[64];

// returns SG_OK with the mutex HELD (caller unlocks), or an error with it released
int scratch_acquire(int device, size_t bytes, CallScratch **out)
{
    if (device < 0 || device >= 64) return SG_ERR_INVALID;
    CallScratch &c = g_call_scratch[device];
    c.mu.lock();
    cudaError_t e = cudaSuccess;
    if (!c.st) e = cudaStreamCreateWithFlags(&c.st, cudaStreamNonBlocking);
    if (e == cudaSuccess && bytes > c.cap) {
        if (c.d) cudaFree(c.d);
        if (c.h) cudaFreeHost(c.h);
        c.d = c.h = nullptr;
        c.cap = 0;
        const size_t want = bytes < (size_t)1 << 16 ? (size_t)1 << 16 : bytes;
        e = cudaMalloc(&c.d, want);
        if (e == cudaSuccess) e = cudaHostAlloc(&c.h, want, cudaHostAllocDefault);
        if (e == cudaSuccess) c.cap = want;
    }
    if (e != cudaSuccess) {
        if (c.d) cudaFree(c.d);
        if (c.h) cudaFreeHost(c.h);
        c.d = c.h = nullptr;
        c.cap = 0;
        c.mu.unlock();
        set_error("scratch allocation: %s", cudaGetErrorString(e));
        return SG_ERR_OOM;
    }
    *out = &c;
    return SG_OK;
}
}  // namespace

extern "C" {

int sg_fuse_topk(int device, const sg_fuse_weights *w, const sg_fuse_config *cfg, const sg_candidate_feats *f,
                 uint32_t n, int64_t now_unix_s, uint32_t *order_out, float *score_out, uint32_t *n_out,
                 float *all_scores_out)
{
    if (!w || !cfg || !f || !n_out || (cfg->max_results && (!order_out || !score_out))) return SG_ERR_INVALID;
    *n_out = 0;
    if (n == 0 || cfg->max_results == 0) return SG_OK;
    if (n > 8192) {
        set_error("sg_fuse_topk supports up to 8192 candidates per call (got %u)", n);
        return SG_ERR_UNSUPPORTED;
    }
    int rc = check_device(device, nullptr);
    if (rc != SG_OK) return rc;
    SG_CUDA_TRY(cudaSetDevice(device));
    // one transient arena: 6 f32 + 1 u32 + 1 i64 per candidate in, order/score/all/n out
    const size_t in_bytes = (size_t)n * (7 * 4 + 8);
    const size_t out_bytes = (size_t)cfg->max_results * 8 + (size_t)n * 4 + 16;
    CallScratch *sc = nullptr;
    rc = scratch_acquire(device, in_bytes + out_bytes + 32, &sc);
    if (rc != SG_OK) return rc;
    uint8_t *d = sc->d, *h = sc->h;
    cudaStream_t st = sc->st;
    cudaError_t e = cudaSuccess;
    int ret = SG_OK;
    {
        // created_at first (8-byte aligned), then the 4-byte arrays
        size_t off = 0;
        memcpy(h + off, f->created_at_unix_s, (size_t)n * 8); const size_t o_created = off; off += (size_t)n * 8;
        const void *srcs[7] = {f->semantic, f->entity, f->tag, f->importance, f->momentum, f->graph_strength, f->access_count};
        size_t offs[7];
        for (int i = 0; i < 7; ++i) {
            memcpy(h + off, srcs[i], (size_t)n * 4);
            offs[i] = off;
            off += (size_t)n * 4;
        }
        const size_t o_n = (off + 15) & ~(size_t)15;
        const size_t o_order = o_n + 16, o_score = o_order + (size_t)cfg->max_results * 4;
        const size_t o_all = o_score + (size_t)cfg->max_results * 4;
        e = cudaMemcpyAsync(d, h, in_bytes, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) {
            FuseDev fd;
            fd.created = reinterpret_cast<const int64_t *>(d + o_created);
            fd.semantic = reinterpret_cast<const float *>(d + offs[0]);
            fd.entity = reinterpret_cast<const float *>(d + offs[1]);
            fd.tag = reinterpret_cast<const float *>(d + offs[2]);
            fd.importance = reinterpret_cast<const float *>(d + offs[3]);
            fd.momentum = reinterpret_cast<const float *>(d + offs[4]);
            fd.graph = reinterpret_cast<const float *>(d + offs[5]);
            fd.access = reinterpret_cast<const uint32_t *>(d + offs[6]);
            e = launch_fuse_rank(fd, n, *w, *cfg, now_unix_s, reinterpret_cast<uint32_t *>(d + o_order),
                                 reinterpret_cast<float *>(d + o_score), reinterpret_cast<uint32_t *>(d + o_n),
                                 reinterpret_cast<float *>(d + o_all), st);
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(h + o_n, d + o_n, o_all + (size_t)n * 4 - o_n, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e == cudaSuccess) {
            uint32_t m = *reinterpret_cast<uint32_t *>(h + o_n);
            *n_out = m;
            memcpy(order_out, h + o_order, (size_t)m * 4);
            memcpy(score_out, h + o_score, (size_t)m * 4);
            if (all_scores_out) memcpy(all_scores_out, h + o_all, (size_t)n * 4);
        }
    }
    if (e != cudaSuccess) {
        set_error("sg_fuse_topk: %s", cudaGetErrorString(e));
        ret = SG_ERR_CUDA;
    }
    sc->mu.unlock();
    return ret;
}

int sg_rrf_fuse(int device, float k, const float *weights, uint32_t n_weights, const uint64_t *mem_ids,
                const uint32_t *list_offsets, uint32_t n_lists, uint64_t *mem_out, float *score_out, uint32_t *n_out)
{
    if (!n_out || !list_offsets || (n_weights && !weights)) return SG_ERR_INVALID;
    *n_out = 0;
    if (n_lists == 0) return SG_OK;
    const uint32_t n_total = list_offsets[n_lists];
    for (uint32_t l = 0; l < n_lists; ++l)
        if (list_offsets[l] > list_offsets[l + 1]) {
            set_error("sg_rrf_fuse: list_offsets must be non-decreasing");
            return SG_ERR_INVALID;
        }
    if (n_total == 0) return SG_OK;
    if (!mem_ids || !mem_out || !score_out) return SG_ERR_INVALID;
    if (n_total > 2048 || n_weights > 64 || n_lists > 64) {
        set_error("sg_rrf_fuse supports up to 2048 ranked entries and 64 lists per call (got %u entries, %u lists)", n_total, n_lists);
        return SG_ERR_UNSUPPORTED;
    }
    int rc = check_device(device, nullptr);
    if (rc != SG_OK) return rc;
    SG_CUDA_TRY(cudaSetDevice(device));
    // arena: ids [n,2] u64 | offsets | weights || ids_out [n,2] u64 | scores | n_out
    const size_t o_ids = 0, o_off = (size_t)n_total * 16, o_w = o_off + ((size_t)(n_lists + 1) * 4 + 15) / 16 * 16;
    const size_t in_bytes = o_w + ((size_t)(n_weights ? n_weights : 1) * 4 + 15) / 16 * 16;
    const size_t o_out = in_bytes, o_sc = o_out + (size_t)n_total * 16, o_n = o_sc + ((size_t)n_total * 4 + 15) / 16 * 16;
    const size_t total = o_n + 16;
    CallScratch *sc = nullptr;
    rc = scratch_acquire(device, total, &sc);
    if (rc != SG_OK) return rc;
    uint8_t *d = sc->d, *h = sc->h;
    cudaStream_t st = sc->st;
    cudaError_t e = cudaSuccess;
    int ret = SG_OK;
    {
        memcpy(h + o_ids, mem_ids, (size_t)n_total * 16);
        memcpy(h + o_off, list_offsets, (size_t)(n_lists + 1) * 4);
        if (n_weights) memcpy(h + o_w, weights, (size_t)n_weights * 4);
        e = cudaMemcpyAsync(d, h, in_bytes, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess)
            e = launch_rrf_fuse(k, reinterpret_cast<const float *>(d + o_w), n_weights, reinterpret_cast<const uint64_t *>(d + o_ids),
                                reinterpret_cast<const uint32_t *>(d + o_off), n_lists, n_total,
                                reinterpret_cast<uint64_t *>(d + o_out), reinterpret_cast<float *>(d + o_sc),
                                reinterpret_cast<uint32_t *>(d + o_n), st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h + o_out, d + o_out, total - o_out, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e == cudaSuccess) {
            const uint32_t m = *reinterpret_cast<uint32_t *>(h + o_n);
            *n_out = m;
            memcpy(mem_out, h + o_out, (size_t)m * 16);
            memcpy(score_out, h + o_sc, (size_t)m * 4);
        }
    }
    if (e != cudaSuccess) {
        set_error("sg_rrf_fuse: %s", cudaGetErrorString(e));
        ret = SG_ERR_CUDA;
    }
    sc->mu.unlock();
    return ret;
}

}  // extern "C"
