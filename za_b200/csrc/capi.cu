// extern "C" surface of libza_b200.so (declared in include/za_b200.h).
// Every entry point converts C++ exceptions into status codes; nothing crosses the ABI but plain
// pointers and sizes.
#include "common.cuh"
#include <stdlib.h>
#include "api_internal.cuh"
#include <string.h>

namespace za {

std::string& last_error() {
    static thread_local std::string s;
    return s;
}
int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    last_error() = buf;
    return code;
}
Ctx::~Ctx() {
    for (auto& kv : domains) delete kv.second;
    for (MsmSlot& sl : slots) {
        if (sl.host_win) cudaFreeHost(sl.host_win);
        if (sl.acc_done) cudaEventDestroy(sl.acc_done);
        if (sl.done) cudaEventDestroy(sl.done);
        if (sl.sort_ev) cudaEventDestroy(sl.sort_ev);
        if (sl.side) cudaStreamDestroy(sl.side);
        if (sl.side2) cudaStreamDestroy(sl.side2);
        if (sl.red_fork) cudaEventDestroy(sl.red_fork);
        if (sl.red_join) cudaEventDestroy(sl.red_join);
    }
    if (side) cudaStreamDestroy(side);
    if (g2_stream) cudaStreamDestroy(g2_stream);
    if (h_stream) cudaStreamDestroy(h_stream);
    if (h_fork) cudaEventDestroy(h_fork);
    if (h_join) cudaEventDestroy(h_join);
    if (g2_fork) cudaEventDestroy(g2_fork);
    if (host_flag) cudaFreeHost(host_flag);
    if (own_stream && stream) cudaStreamDestroy(stream);
}

}  // namespace za

using namespace za;


#define ZA_TRY try {
#define ZA_CATCH                                                                   \
    }                                                                              \
    catch (const ZaError& e) { return fail(e.code, "%s", e.what()); }              \
    catch (const CudaError& e) { return fail(ZA_ERR_CUDA, "%s", e.what()); }       \
    catch (const std::bad_alloc&) { return fail(ZA_ERR_INVALID, "out of host memory"); } \
    catch (const std::exception& e) { return fail(ZA_ERR_INVALID, "%s", e.what()); }

extern "C" {

const char* za_last_error(void) { return last_error().c_str(); }
int za_version(void) { return 1; }

int za_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int za_ctx_create(int device, za_ctx** out) {
    if (!out) return fail(ZA_ERR_INVALID, "za_ctx_create: out is NULL");
    *out = nullptr;
    ZA_TRY
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(ZA_ERR_CUDA, "no CUDA device available (%s); za_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= n) return fail(ZA_ERR_INVALID, "device %d out of range (have %d)", device, n);
    ZA_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    ZA_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(ZA_ERR_CUDA, "device %d is sm_%d%d; za_b200 is built for sm_100a only", device, prop.major, prop.minor);
    za_ctx* c = new za_ctx();
    c->c.device = device;
    c->c.sm_count = prop.multiProcessorCount;
    {
        // The multiexps gather 64-byte (G1) / 128-byte (G2) points at random from tables far larger than L2; the
        // default L2 fill granularity fetches 128 bytes per miss.  ZA_L2_FETCH = 32 / 64 / 128 overrides.
        size_t gran = 0;
        if (const char* e = getenv("ZA_L2_FETCH")) gran = (size_t)atoi(e);
        if (gran) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
        size_t got = 0;
        cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
        if (getenv("ZA_DEBUG_TIMELINE")) fprintf(stderr, "[za] L2 fetch granularity %zu\n", got);
        cudaGetLastError();
    }
    ZA_CUDA(cudaStreamCreateWithFlags(&c->own, cudaStreamNonBlocking));
    ZA_CUDA(cudaStreamCreateWithFlags(&c->c.side, cudaStreamNonBlocking));
    c->c.stream = c->own;
    c->c.own_stream = false;  // destroyed below, not by ~Ctx
    *out = c;
    return ZA_OK;
    ZA_CATCH
}

void za_ctx_destroy(za_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->c.device);
    cudaStreamSynchronize(ctx->c.stream);
    if (ctx->c.side) cudaStreamSynchronize(ctx->c.side);
    if (ctx->own) cudaStreamDestroy(ctx->own);
    delete ctx;
}

int za_ctx_set_stream(za_ctx* ctx, void* cuda_stream) {
    if (!ctx) return fail(ZA_ERR_INVALID, "ctx is NULL");
    ctx->c.stream = (cudaStream_t)cuda_stream;      // NULL is the CUDA default stream
    return ZA_OK;
}

int za_ctx_use_own_stream(za_ctx* ctx) {
    if (!ctx) return fail(ZA_ERR_INVALID, "ctx is NULL");
    ctx->c.stream = ctx->own;
    return ZA_OK;
}

int za_ctx_synchronize(za_ctx* ctx) {
    if (!ctx) return fail(ZA_ERR_INVALID, "ctx is NULL");
    ZA_TRY
    ZA_CUDA(cudaSetDevice(ctx->c.device));
    ZA_CUDA(cudaStreamSynchronize(ctx->c.stream));
    return ZA_OK;
    ZA_CATCH
}

uint64_t za_ctx_launch_count(const za_ctx* ctx) { return ctx ? ctx->c.launches : 0; }

int za_ctx_profile(za_ctx* ctx, int on) {
    if (!ctx) return fail(ZA_ERR_INVALID, "ctx is NULL");
    ctx->c.profile = on != 0;
    return ZA_OK;
}

int za_ctx_profile_read(za_ctx* ctx, double* out) {
    if (!ctx || !out) return fail(ZA_ERR_INVALID, "NULL argument");
    ZA_TRY
    Ctx* c = &ctx->c;
    ZA_CUDA(cudaSetDevice(c->device));
    ZA_CUDA(cudaStreamSynchronize(c->stream));
    double ms[PROF_NCAT] = {0};
    for (ProfSpan& sp : c->spans) {
        float t = 0;
        if (cudaEventElapsedTime(&t, sp.a, sp.b) == cudaSuccess) ms[sp.cat] += t;
        cudaEventDestroy(sp.a); cudaEventDestroy(sp.b);
    }
    c->spans.clear();
    for (int i = 0; i < PROF_NCAT; i++) {
        out[i] = ms[i]; out[PROF_NCAT + i] = c->prof_work[i]; out[2 * PROF_NCAT + i] = (double)c->prof_count[i];
        c->prof_work[i] = 0; c->prof_count[i] = 0;
    }
    return ZA_OK;
    ZA_CATCH
}

// ------------------------------------------------------------------------------------- NTT
int za_fr_convert_device(za_ctx* ctx, void* d_data, size_t n, int dir) {
    if (!ctx || !d_data) return fail(ZA_ERR_INVALID, "NULL argument");
    ZA_TRY
    ZA_CUDA(cudaSetDevice(ctx->c.device));
    fr_convert(&ctx->c, (Fr*)d_data, n, dir);
    return ZA_OK;
    ZA_CATCH
}

int za_ntt_device(za_ctx* ctx, void* d_data, int log_n, int mode, int batch) {
    if (!ctx || !d_data) return fail(ZA_ERR_INVALID, "NULL argument");
    if (log_n < 0 || batch < 1 || mode < 0 || mode > 3) return fail(ZA_ERR_INVALID, "bad log_n/mode/batch");
    ZA_TRY
    ZA_CUDA(cudaSetDevice(ctx->c.device));
    ntt_mode(&ctx->c, (Fr*)d_data, log_n, mode, batch);
    return ZA_OK;
    ZA_CATCH
}

static void check_canonical_host(const uint8_t* p, size_t n, const char* what) {
    for (size_t i = 0; i < n; i++) {
        uint32_t w[8];
        memcpy(w, p + 32 * i, 32);
        if (!fp_is_canonical<FrParams>(w)) {
            char b[128];
            snprintf(b, sizeof b, "%s[%zu] is not a canonical Fr element (>= r)", what, i);
            throw ZaError(ZA_ERR_NOT_CANONICAL, b);
        }
    }
}

int za_ntt(za_ctx* ctx, uint8_t* data, int log_n, int mode) {
    if (!ctx || !data) return fail(ZA_ERR_INVALID, "NULL argument");
    if (log_n < 0 || mode < 0 || mode > 3) return fail(ZA_ERR_INVALID, "bad log_n/mode");
    ZA_TRY
    if (log_n >= 28) throw ZaError(ZA_ERR_POLY_DEGREE_TOO_LARGE, "domain of 2^28 or more elements");
    Ctx* c = &ctx->c;
    ZA_CUDA(cudaSetDevice(c->device));
    size_t N = (size_t)1 << log_n;
    check_canonical_host(data, N, "data");
    DevBuf& buf = c->scratch[1];
    buf.ensure(N * sizeof(Fr));
    ZA_CUDA(cudaMemcpyAsync(buf.p, data, N * 32, cudaMemcpyHostToDevice, c->stream));
    fr_convert(c, buf.as<Fr>(), N, 0);
    ntt_mode(c, buf.as<Fr>(), log_n, mode, 1);
    fr_convert(c, buf.as<Fr>(), N, 1);
    ZA_CUDA(cudaMemcpyAsync(data, buf.p, N * 32, cudaMemcpyDeviceToHost, c->stream));
    ZA_CUDA(cudaStreamSynchronize(c->stream));
    return ZA_OK;
    ZA_CATCH
}

int za_h_poly_device(za_ctx* ctx, void* d_a, void* d_b, void* d_c, int log_m) {
    if (!ctx || !d_a || !d_b || !d_c) return fail(ZA_ERR_INVALID, "NULL argument");
    ZA_TRY
    ZA_CUDA(cudaSetDevice(ctx->c.device));
    h_poly_device(&ctx->c, (Fr*)d_a, (Fr*)d_b, (Fr*)d_c, log_m);
    return ZA_OK;
    ZA_CATCH
}

int za_h_poly(za_ctx* ctx, const uint8_t* a, const uint8_t* b, const uint8_t* c, size_t len, uint8_t* h_out, uint8_t* checkpoints) {
    if (!ctx || !a || !b || !c || !h_out) return fail(ZA_ERR_INVALID, "NULL argument");
    ZA_TRY
    Ctx* cx = &ctx->c;
    ZA_CUDA(cudaSetDevice(cx->device));
    size_t m = 1;
    int log_m = 0;
    while (m < len) { m *= 2; log_m++; if (log_m >= 28) throw ZaError(ZA_ERR_POLY_DEGREE_TOO_LARGE, "domain of 2^28 or more elements"); }
    check_canonical_host(a, len, "a"); check_canonical_host(b, len, "b"); check_canonical_host(c, len, "c");
    DevBuf& buf = cx->scratch[1];
    buf.ensure(3 * m * sizeof(Fr));
    Fr* d = buf.as<Fr>();
    ZA_CUDA(cudaMemsetAsync(d, 0, 3 * m * sizeof(Fr), cx->stream));
    const uint8_t* src[3] = {a, b, c};
    for (int v = 0; v < 3; v++) ZA_CUDA(cudaMemcpyAsync(d + v * m, src[v], len * 32, cudaMemcpyHostToDevice, cx->stream));
    fr_convert(cx, d, 3 * m, 0);
    if (checkpoints) h_poly_checkpointed(cx, d, d + m, d + 2 * m, log_m, checkpoints);
    else h_poly_device(cx, d, d + m, d + 2 * m, log_m);
    if (m > 1) ZA_CUDA(cudaMemcpyAsync(h_out, d, (m - 1) * 32, cudaMemcpyDeviceToHost, cx->stream));
    ZA_CUDA(cudaStreamSynchronize(cx->stream));
    return ZA_OK;
    ZA_CATCH
}

}  // extern "C"
