// Shared host-side plumbing for the za_b200 CUDA library: error reporting and the context object.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <map>
#include <stdexcept>
#include "../../include/za_b200.h"
#include "ec.cuh"

namespace za {

// thread-local last-error text, returned by za_last_error()
std::string& last_error();
int fail(int code, const char* fmt, ...);

struct CudaError : std::runtime_error {
    explicit CudaError(const std::string& s) : std::runtime_error(s) {}
};
struct ZaError : std::runtime_error {
    int code;
    ZaError(int c, const std::string& s) : std::runtime_error(s), code(c) {}
};

#define ZA_CUDA(expr)                                                                                   \
    do {                                                                                                \
        cudaError_t e_ = (expr);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            char b_[512];                                                                               \
            snprintf(b_, sizeof b_, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); \
            throw za::CudaError(b_);                                                                    \
        }                                                                                               \
    } while (0)

// RAII device buffer (cudaMallocAsync on the context stream would need pool tuning; plain cudaMalloc
// is fine because every long-lived buffer is cached in the context or the proving key).
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    DevBuf() {}
    explicit DevBuf(size_t n) { alloc(n); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), bytes(o.bytes) { o.p = nullptr; o.bytes = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; bytes = o.bytes; o.p = nullptr; o.bytes = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t n) {
        release();
        if (n == 0) n = 16;
        ZA_CUDA(cudaMalloc(&p, n));
        bytes = n;
    }
    void ensure(size_t n) { if (n > bytes) alloc(n); }
    void release() { if (p) { cudaFree(p); p = nullptr; bytes = 0; } }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Per-domain-size tables (bellman's EvaluationDomain fields: omega, omegainv, geninv, minv —
// here expanded into device tables once and cached).
struct NttDomain {
    int log_n = 0;
    DevBuf tw;          // omega^e, e in [0, N/2]    (N/2 + 1 entries; the last one is -1)
    DevBuf pow_g;       // g^i                      (coset_fft pre-scale), built lazily
    DevBuf pow_g_minv;  // m^-1 g^i                 (fused ifft -> coset_fft)
    DevBuf pow_ginv_minv;  // m^-1 g^-i             (icoset_fft post-scale)
    DevBuf pow_ginv_minv_canon;  // the same out of Montgomery form: the last store of the H pipeline leaves canonical values
    DevBuf consts;      // [0] = m^-1, [1] = (g^m - 1)^-1
    bool have_coset = false;
};

// Per-multiexp working set, so that several multiexps can be in flight (sort + accumulate on the main
// stream, bucket reduction on the side stream) and their window sums collected at the end.
struct MsmSlot {
    DevBuf counts, entries, buckets, parts, segs;
    DevBuf red;                          // row / column bucket reduction: stage buffers, row sums, six results per bucket space
    int red_mode = 0, red_k = 0;         // 1: row / column reduction (K6'), host_win holds 6 points per bucket space
    uint32_t red_H = 0, red_Lw = 0;
    int red_sR = 0, red_sC = 0;          // K6'c: the row / column arrays are split once more; the host scales the outer part by 2^shift
    DevBuf pair_pts[2], pair_offs;       // batched-affine pair rounds: ping-pong point buffers, per-round bucket offsets
    int rounds = 0;
    void* host_win = nullptr;            // pinned: window sums (+ entry count when profiling)
    size_t host_win_bytes = 0;
    cudaEvent_t acc_done = nullptr, done = nullptr;
    cudaEvent_t sort_ev = nullptr;        // recorded behind this slot's digit sort
    cudaStream_t side = nullptr;         // this slot's reduction stream (high priority, so it is not starved by accumulations)
    cudaStream_t side2 = nullptr;        // row sums of the row / column reduction run here, next to the column sums on `side`
    cudaEvent_t red_fork = nullptr, red_join = nullptr;
    cudaEvent_t dbg_start = nullptr, dbg_sort = nullptr, dbg_acc = nullptr, dbg_done = nullptr;   // ZA_DEBUG_TIMELINE only
    bool busy = false;                   // enqueued, not yet finished
    bool done_valid = false;             // `done` has been recorded at least once
    uint32_t sort_users = 0;             // slots that reuse this slot's digit sort since its last enqueue
    int kind = 0;                        // 0 empty, 1 naive (small), 2 bucket method
    int c = 0, W = 0, warps = 0, acc_cat = 0;
    int nparts = 1;                      // queries sorted and accumulated together in this slot (K4')
    bool single = false;                 // fixed-base table mode: one bucket space per query
    uint32_t nkeys = 0, Lc = 0, nchunks = 0;
    size_t n = 0;
    uint32_t* d_offsets = nullptr;       // sort result (may be shared by a later multiexp over the same scalars)
    uint32_t* d_entries = nullptr;
};

enum { PROF_ACC_G1 = 0, PROF_ACC_G2 = 1, PROF_NTT = 2, PROF_MSM_SORT = 3, PROF_MSM_REDUCE = 4, PROF_R1CS = 5, PROF_POINTWISE = 6, PROF_OTHER = 7, PROF_NCAT = 8 };
struct ProfSpan { cudaEvent_t a, b; int cat; };

// Where the last store of the H pipeline puts h[k]: the first part j with k < hi[j] receives it at out[j] + k (out[j] may be
// memory of a peer device — the h slices of a multi-GPU proof travel as the stores of the last NTT pass, not as copies).
#define ZA_H_SCATTER_MAX 16
struct HScatter {
    int n = 0;
    Fr* out[ZA_H_SCATTER_MAX];
    uint32_t hi[ZA_H_SCATTER_MAX];
};

struct Ctx {
    int device = 0;
    cudaStream_t stream = nullptr;      // stream all work of this context is issued on
    bool own_stream = false;
    int sm_count = 0;
    std::map<int, NttDomain*> domains;  // by log_n
    // scratch reused across calls
    DevBuf scratch[16];
    MsmSlot slots[8];
    cudaStream_t side = nullptr;        // bucket reductions overlap the next multiexp's accumulation here
    cudaStream_t g2_stream = nullptr;   // the G2 multiexp of a proof runs here, next to the G1 multiexps (prove_msms_enqueue)
    cudaEvent_t g2_fork = nullptr;
    cudaStream_t h_stream = nullptr;    // the H pipeline runs here, next to the witness multiexps (create_proof_device)
    cudaEvent_t h_fork = nullptr, h_join = nullptr;
    unsigned long long* host_flag = nullptr;   // pinned: verdict of the witness range check of create_proof
    uint64_t launches = 0;              // kernels launched through this context (bench: gpu_launches)
    cudaEvent_t dbg_t0 = nullptr;       // ZA_DEBUG_TIMELINE: start of the current proof on this device
    bool dbg_t0_valid = false;
    bool witness_merged = false;        // the last prove enqueued B (G1), L and A as one multiexp in slot 3
    HScatter h_scatter;                 // n > 0: destinations of the next H pipeline's output (prove_h consumes and clears it)
    bool ntt_attr_set = false;          // the > 48 KiB shared-memory attribute of the NTT kernels is set on this context's device
    // optional per-kernel-class timing with CUDA events on `stream` (bench.py roofline numbers)
    bool profile = false;
    std::vector<ProfSpan> spans;
    double prof_work[PROF_NCAT] = {0};  // algorithmic work units per class (mixed additions, elements, ...)
    uint64_t prof_count[PROF_NCAT] = {0};
    ~Ctx();
};

NttDomain* get_domain(Ctx* ctx, int log_n, bool need_coset);

// RAII: brackets the kernels issued in its scope with two events when profiling is on
struct ProfScope {
    Ctx* c; int cat; cudaEvent_t a = nullptr, b = nullptr;
    cudaStream_t st;
    ProfScope(Ctx* ctx, int category, double work = 0, cudaStream_t on = nullptr, bool use_on = false) : c(ctx), cat(category) {
        st = use_on ? on : c->stream;
        if (!c->profile) return;
        cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a, st);
        c->prof_work[cat] += work; c->prof_count[cat]++;
    }
    ~ProfScope() {
        if (!a) return;
        cudaEventRecord(b, st);
        c->spans.push_back(ProfSpan{a, b, cat});
    }
};

}  // namespace za
