#include "AMReX_Parallel.H"
#include "AMReX_Base.H"

#include <cuda_runtime.h>
#include <nccl.h>

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <unordered_map>
#include <string>
#include <vector>
#include <algorithm>

namespace amrex {

namespace {
    int g_rank = 0, g_nranks = 1;
    ncclComm_t g_comm = nullptr;
    bool g_gpu_init = false;
    int g_device = -1;
    cudaStream_t g_stream = nullptr, g_comm_stream = nullptr, g_aux_stream = nullptr, g_override = nullptr;
    long long g_launches = 0;
    double* g_red_dev = nullptr;     // staging for scalar all-reduces
    double* g_red_pin = nullptr;
    constexpr int kRedMax = 64;

    void nccl_check (ncclResult_t r, const char* what)
    {
        if (r != ncclSuccess) { Abort(std::string("NCCL failure in ") + what + ": " + ncclGetErrorString(r)); }
    }
}

namespace Gpu {

void check (int err, const char* what, const char* file, int line)
{
    if (err != int(cudaSuccess)) {
        Abort(std::string("CUDA error ") + cudaGetErrorString(cudaError_t(err)) + " in " + what + " at " + file + ":" + std::to_string(line));
    }
}

void Initialize (int device_id)
{
    if (g_gpu_init) { return; }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        Abort("amrex_b200: no CUDA device visible - this library has no CPU fallback");
    }
    g_device = device_id % n;
    AMREX_CUDA_SAFE_CALL(cudaSetDevice(g_device));
    AMREX_CUDA_SAFE_CALL(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
    AMREX_CUDA_SAFE_CALL(cudaStreamCreateWithFlags(&g_comm_stream, cudaStreamNonBlocking));
    AMREX_CUDA_SAFE_CALL(cudaStreamCreateWithFlags(&g_aux_stream, cudaStreamNonBlocking));
    AMREX_CUDA_SAFE_CALL(cudaMalloc(&g_red_dev, kRedMax * sizeof(double)));
    AMREX_CUDA_SAFE_CALL(cudaMallocHost(&g_red_pin, kRedMax * sizeof(double)));
    g_gpu_init = true;
}

void Finalize ()
{
    if (!g_gpu_init) { return; }
    cudaDeviceSynchronize();
    The_Arena()->freeUnused();
    cudaFree(g_red_dev); cudaFreeHost(g_red_pin);
    cudaStreamDestroy(g_stream); cudaStreamDestroy(g_comm_stream); cudaStreamDestroy(g_aux_stream);
    g_stream = g_comm_stream = g_aux_stream = nullptr;
    g_gpu_init = false;
}

bool Initialized () noexcept { return g_gpu_init; }
bool debugSync () noexcept { static const bool on = std::getenv("B200MG_DEBUG_SYNC") != nullptr; return on; }
int debugSyncNow () noexcept { return int(cudaDeviceSynchronize()); }
int deviceId () noexcept { return g_device; }
cudaStream_t gpuStream () noexcept { return g_override ? g_override : g_stream; }
cudaStream_t commStream () noexcept { return g_comm_stream; }
cudaStream_t auxStream () noexcept { return g_aux_stream; }
void setStream (cudaStream_t s) noexcept { g_override = s; }
void streamSynchronize () { AMREX_CUDA_SAFE_CALL(cudaStreamSynchronize(gpuStream())); }
void synchronize () { AMREX_CUDA_SAFE_CALL(cudaDeviceSynchronize()); }
void htod_memcpy_async (void* d, const void* s, std::size_t n) { AMREX_CUDA_SAFE_CALL(cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, gpuStream())); }
void dtoh_memcpy_async (void* d, const void* s, std::size_t n) { AMREX_CUDA_SAFE_CALL(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, gpuStream())); }
void dtod_memcpy_async (void* d, const void* s, std::size_t n) { AMREX_CUDA_SAFE_CALL(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToDevice, gpuStream())); }
void memset_async (void* d, int v, std::size_t n) { AMREX_CUDA_SAFE_CALL(cudaMemsetAsync(d, v, n, gpuStream())); }
long long launchCount () noexcept { return g_launches; }
void resetLaunchCount () noexcept { g_launches = 0; }
void countLaunch (int n) noexcept { g_launches += n; }

namespace {
    struct ProfRec { std::string name; int scope; cudaEvent_t e0, e1; };
    bool g_prof_on = false;
    int g_prof_scope = -1;
    std::vector<ProfRec> g_prof;
    std::vector<cudaEvent_t> g_prof_pool;
    cudaEvent_t prof_event ()
    {
        if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
        cudaEvent_t e; AMREX_CUDA_SAFE_CALL(cudaEventCreate(&e)); return e;
    }
}
void profileEnable (bool on) { g_prof_on = on; }
bool profiling () noexcept { return g_prof_on; }
void profileScope (int s) noexcept { g_prof_scope = s; }
int profileScope () noexcept { return g_prof_scope; }
void profileBegin (const char* call_text)
{
    const char* p = std::strchr(call_text, '(');
    ProfRec r{p ? std::string(call_text, p - call_text) : std::string(call_text), g_prof_scope, prof_event(), prof_event()};
    while (!r.name.empty() && r.name.back() == ' ') { r.name.pop_back(); }
    AMREX_CUDA_SAFE_CALL(cudaEventRecord(r.e0, gpuStream()));
    g_prof.push_back(std::move(r));
}
void profileEnd () { AMREX_CUDA_SAFE_CALL(cudaEventRecord(g_prof.back().e1, gpuStream())); }
std::string profileReport ()
{
    AMREX_CUDA_SAFE_CALL(cudaStreamSynchronize(gpuStream()));
    struct Acc { long long n = 0; double tot = 0, mn = 1e30, mx = 0; };
    std::map<std::pair<std::string, int>, Acc> acc;
    for (auto& r : g_prof) {
        float ms = 0.f;
        AMREX_CUDA_SAFE_CALL(cudaEventElapsedTime(&ms, r.e0, r.e1));
        Acc& a = acc[{r.name, r.scope}];
        a.n++; a.tot += ms; a.mn = std::min(a.mn, double(ms)); a.mx = std::max(a.mx, double(ms));
        g_prof_pool.push_back(r.e0); g_prof_pool.push_back(r.e1);
    }
    g_prof.clear();
    std::string out;
    char line[256];
    for (auto const& kv : acc) {
        std::snprintf(line, sizeof(line), "%s %d %lld %.6f %.6f %.6f\n", kv.first.first.c_str(), kv.first.second, kv.second.n,
                      kv.second.tot, kv.second.mn, kv.second.mx);
        out += line;
    }
    return out;
}

} // namespace Gpu

namespace ParallelDescriptor {

int MyProc () noexcept { return g_rank; }
int NProcs () noexcept { return g_nranks; }
int NcclUniqueIdBytes () noexcept { return int(sizeof(ncclUniqueId)); }
void NcclGetUniqueId (void* out) { ncclUniqueId id; nccl_check(ncclGetUniqueId(&id), "ncclGetUniqueId"); std::memcpy(out, &id, sizeof(id)); }

void InitComm (int rank, int nranks, const void* uid)
{
    g_rank = rank; g_nranks = nranks;
    if (nranks > 1) {
        AMREX_ALWAYS_ASSERT_WITH_MESSAGE(Gpu::Initialized(), "Gpu::Initialize must precede InitComm");
        ncclUniqueId id; std::memcpy(&id, uid, sizeof(id));
        nccl_check(ncclCommInitRank(&g_comm, nranks, id, rank), "ncclCommInitRank");
    }
}

void FinalizeComm ()
{
    if (g_comm) { ncclCommDestroy(g_comm); g_comm = nullptr; }
    g_rank = 0; g_nranks = 1;
}

bool HasComm () noexcept { return g_comm != nullptr; }
void* Comm () noexcept { return g_comm; }

namespace {
template <class T>
void reduce_small (T* v, int n, ncclDataType_t dt, ncclRedOp_t op)
{
    if (g_nranks == 1) { return; }
    AMREX_ALWAYS_ASSERT(n * int(sizeof(T)) <= kRedMax * int(sizeof(double)));
    cudaStream_t s = Gpu::gpuStream();
    std::memcpy(g_red_pin, v, n * sizeof(T));
    AMREX_CUDA_SAFE_CALL(cudaMemcpyAsync(g_red_dev, g_red_pin, n * sizeof(T), cudaMemcpyHostToDevice, s));
    nccl_check(ncclAllReduce(g_red_dev, g_red_dev, n, dt, op, g_comm, s), "ncclAllReduce");
    AMREX_CUDA_SAFE_CALL(cudaMemcpyAsync(g_red_pin, g_red_dev, n * sizeof(T), cudaMemcpyDeviceToHost, s));
    AMREX_CUDA_SAFE_CALL(cudaStreamSynchronize(s));
    std::memcpy(v, g_red_pin, n * sizeof(T));
}
}

void ReduceRealSum (double* v, int n) { reduce_small(v, n, ncclDouble, ncclSum); }
void ReduceRealMax (double* v, int n) { reduce_small(v, n, ncclDouble, ncclMax); }
void ReduceRealMin (double* v, int n) { reduce_small(v, n, ncclDouble, ncclMin); }
void ReduceLongSum (long long* v, int n) { reduce_small(v, n, ncclInt64, ncclSum); }
void Barrier () { double z = 0; ReduceRealSum(&z, 1); }

double second () noexcept
{
    static const auto t0 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

} // namespace ParallelDescriptor

// -------------------------------------------------------------------------------------------- Arena
struct Arena::Impl {
    std::mutex mtx;
    std::multimap<std::size_t, void*> free_blocks;
    std::unordered_map<void*, std::size_t> live;
};

namespace { Arena g_arena; Arena::Impl* g_arena_impl = nullptr; }

Arena* The_Arena ()
{
    if (!g_arena_impl) { g_arena_impl = new Arena::Impl; }
    return &g_arena;
}

void* Arena::alloc (std::size_t nbytes)
{
    AMREX_ALWAYS_ASSERT_WITH_MESSAGE(Gpu::Initialized(), "device arena used before amrex::Initialize (no CPU fallback)");
    Impl& I = *g_arena_impl;
    const std::size_t sz = ((nbytes ? nbytes : 1) + 255) / 256 * 256;
    std::lock_guard<std::mutex> lk(I.mtx);
    void* p = nullptr;
    auto it = I.free_blocks.find(sz);
    if (it != I.free_blocks.end()) { p = it->second; I.free_blocks.erase(it); }
    else {
        cudaError_t e = cudaMalloc(&p, sz);
        if (e != cudaSuccess) {   // release the cache and retry once
            for (auto& kv : I.free_blocks) { cudaFree(kv.second); m_reserved -= kv.first; }
            I.free_blocks.clear();
            AMREX_CUDA_SAFE_CALL(cudaMalloc(&p, sz));
        }
        m_reserved += sz;
    }
    I.live[p] = sz; m_in_use += sz;
    return p;
}

void Arena::free (void* p)
{
    if (!p) { return; }
    Impl& I = *g_arena_impl;
    std::lock_guard<std::mutex> lk(I.mtx);
    auto it = I.live.find(p);
    AMREX_ALWAYS_ASSERT(it != I.live.end());
    m_in_use -= it->second;
    I.free_blocks.emplace(it->second, p);
    I.live.erase(it);
}

void Arena::freeUnused ()
{
    if (!g_arena_impl) { return; }
    Impl& I = *g_arena_impl;
    std::lock_guard<std::mutex> lk(I.mtx);
    for (auto& kv : I.free_blocks) { cudaFree(kv.second); m_reserved -= kv.first; }
    I.free_blocks.clear();
}

void* pinned_alloc (std::size_t n) { void* p = nullptr; AMREX_CUDA_SAFE_CALL(cudaMallocHost(&p, n ? n : 1)); return p; }
void pinned_free (void* p) { if (p) { cudaFreeHost(p); } }

} // namespace amrex
