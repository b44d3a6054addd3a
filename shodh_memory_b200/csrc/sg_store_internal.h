// sg_store_internal.h -- the store object and the host helpers shared by store.cu and graph_api.cu.
#pragma once
#include <exception>
#include <mutex>
#include <new>
#include <vector>

#include "sg_internal.h"

using namespace sg;

// ---------------------------------------------------------------------------------------
// the store
// ---------------------------------------------------------------------------------------
struct sg_store {
    int device = 0;
    int sm_count = 148;
    uint32_t dim = 0;
    int dtype = SG_DTYPE_F16;
    int metric = SG_METRIC_NEG_DOT;
    int order = SG_ORDER_SCALAR;
    uint32_t id_base = 0;
    uint32_t esize = 2;

    void *d_rows = nullptr;
    uint64_t n = 0, cap = 0;
    // in-place growth (CUDA virtual memory management): d_rows is a reserved VA range, physical chunks are mapped at
    // its end as the store grows -- no second copy of the corpus, no D2D copy, the base address never changes
    bool vmm = false;
    uint64_t va_bytes = 0, mapped_bytes = 0;
    std::vector<unsigned long long> vmm_handles;   // CUmemGenericAllocationHandle of every mapped chunk
    std::vector<uint64_t> vmm_sizes;
    uint32_t *d_deleted = nullptr;       // bitmap, cap/32 (+pad) words
    std::vector<uint32_t> h_deleted;     // host mirror
    uint64_t n_deleted = 0;

    cudaStream_t stream = nullptr;
    mutable std::mutex mu;

    // scratch
    float *d_queries = nullptr;  size_t q_cap = 0;      // floats
    uint64_t *d_lists = nullptr; size_t lists_cap = 0;  // keys
    uint64_t *d_lists2 = nullptr; size_t lists2_cap = 0; // first-level merge output
    uint32_t *d_ids = nullptr; float *d_dist = nullptr; uint32_t *d_counts = nullptr; size_t res_cap = 0, cnt_cap = 0;
    uint32_t *d_resblk = nullptr; size_t resblk_cap = 0;   // sg_store_search: ids | distances | counts in one block
    void *h_pinned = nullptr; size_t pinned_cap = 0;
    float *d_stage = nullptr; size_t stage_cap = 0;     // floats, append staging

    // tensor-core path scratch
    float *d_max_norm = nullptr;                           // device scalar: bound on |row|_2 over the store
    size_t qmeta_cap = 0;                                  // capacity (queries) of gtau / flags / flag_list / hist
    float *d_maxima = nullptr; size_t maxima_cap = 0;      // [n_lists * J][b_pad] sample maxima
    uint32_t *d_gtau = nullptr, *d_flags = nullptr, *d_nflag = nullptr, *d_flag_list = nullptr;
    uint32_t *d_hist = nullptr;                            // [b_pad][32] per-query score histogram of the sweep in flight
    uint32_t *d_arrive = nullptr; uint32_t arrive_total = 0; // monotonic arrival counter of the sweeps' meeting point + its host mirror
    uint32_t *d_cnts = nullptr; size_t cnts_cap = 0;       // [b_pad][n_lists] candidates appended per region
    int tc_mode = 0;                                       // 0 = automatic, 1 = never use the tensor-core path
    cudaEvent_t ev_chain = nullptr;                        // end of the last asynchronous call
    cudaStream_t chain_stream = nullptr; bool chain_pending = false;
    unsigned long long *d_flag_total = nullptr;            // cumulative count of uncertified queries
    uint64_t *d_cand = nullptr; size_t cand_cap_q = 0;     // [b_pad][n_lists][cap_l] candidate regions
    uint64_t tc_searches = 0, tc_flagged = 0;              // statistics (sg_store_tc_stats)

    // vector id -> MemoryId table (IdMapping of RetrievalEngine, retrieval.rs:64-140) + search_ids scratch
    uint64_t *d_vec2mem = nullptr; uint64_t n_map = 0; size_t map_cap = 0;
    uint64_t *d_mem_out = nullptr; float *d_sim_out = nullptr; uint32_t *d_memcnt = nullptr; size_t post_cap = 0, postc_cap = 0;

    // Vamana graph (vamana.cu / graph_api.cu): adjacency [cap][gs] u32, 0xFFFFFFFF-padded; graph_len nodes have lists
    uint32_t *d_graph = nullptr; uint64_t graph_cap = 0;   // rows allocated
    uint32_t gs = 0;                                       // max_degree R (list stride), 0 = no graph
    uint64_t graph_len = 0;
    uint32_t medoid = 0;
    uint32_t search_list_size = 75; float alpha = 1.2f;    // VamanaConfig::default (vamana.rs:80-91)
    uint32_t *d_vis = nullptr; size_t vis_cap = 0;         // visited hash tables of the batch in flight
    uint32_t *d_ann_flags = nullptr; size_t ann_flags_cap = 0;
    uint64_t *d_ann_stats = nullptr; size_t ann_stats_cap = 0;
    uint32_t *d_g32 = nullptr; size_t g32_cap = 0;         // build / insert scratch (u32)
    float *d_gf32 = nullptr; size_t gf32_cap = 0;          // build / insert scratch (f32)
    uint64_t ann_searches = 0, ann_retries = 0;
    unsigned long long *d_ann_phase = nullptr;             // profiling hook: cycles per phase of the beam search

    bool profiling = false;
    std::vector<cudaEvent_t> ev;  // pairs: ev[2i] before, ev[2i+1] after the sweep launches of search i
    uint32_t ev_count = 0;        // searches recorded since the last read/reset
    uint32_t last_scan_launches = 0;
};


namespace sg {

// C ABI bodies that allocate host memory run inside this: a C++ exception must never cross the extern "C" boundary
template <typename F>
int abi_guard(const char *what, F &&body)
{
    try {
        return body();
    } catch (const std::bad_alloc &) {
        set_error("%s: out of host memory", what);
        return SG_ERR_OOM;
    } catch (const std::exception &e) {
        set_error("%s: %s", what, e.what());
        return SG_ERR_INVALID;
    } catch (...) {
        set_error("%s: unexpected exception", what);
        return SG_ERR_INVALID;
    }
}

// forget every row, deleted flag and graph edge (capacity is kept): roll-back of a failed load
int store_truncate_locked(sg_store *s);
int check_device(int device, int *sm_count);
int grow_pinned(sg_store *s, size_t bytes);
int ensure_rows(sg_store *s, uint64_t need_rows);
void vmm_release(sg_store *s);
// exact path A; all pointers are device pointers; asynchronous on st.  use_deleted = false scores every row.
int search_exact_locked(sg_store *s, const float *d_q, uint32_t b, uint32_t k, uint32_t *d_ids, float *d_dist,
                        uint32_t *d_counts, cudaStream_t st, bool record, bool use_deleted = true,
                        const PushArgs *push = nullptr);
// exchange.cu: wait for every rank's pushed lists and merge them (one launch)
int exchange_wait_merge(int device, const PushArgs &p, uint32_t *d_ids_out, float *d_dist_out, uint32_t *d_counts_out,
                        cudaStream_t st);
// Every call that launches work for a store goes through these two: an asynchronous call on stream A followed by a
// call on stream B shares the store's scratch buffers, so B first waits for the event A recorded at its end.
int chain_begin(sg_store *s, cudaStream_t st);
int chain_end(sg_store *s, cudaStream_t st);
// graph_api.cu: link rows [first, first+n) into the graph one by one (VamanaIndex::add_vector); no-op without a graph
int graph_append_locked(sg_store *s, uint64_t first, uint64_t n);
void graph_free(sg_store *s);

template <typename T>
int grow(T *&ptr, size_t &cap, size_t need, bool keep = false, cudaStream_t st = nullptr)
{
    if (need <= cap) return SG_OK;
    size_t ncap = need + need / 2;
    T *np = nullptr;
    SG_CUDA_TRY(cudaMalloc(&np, ncap * sizeof(T)));
    if (keep && ptr && cap) {
        cudaError_t e = cudaMemcpyAsync(np, ptr, cap * sizeof(T), cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            cudaFree(np);
            set_error("device copy during growth failed: %s", cudaGetErrorString(e));
            return SG_ERR_CUDA;
        }
    }
    if (ptr) cudaFree(ptr);
    ptr = np;
    cap = ncap;
    return SG_OK;
}

}  // namespace sg
