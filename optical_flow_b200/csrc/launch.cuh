// launch.cuh -- launch bookkeeping: every kernel launch of the engine goes through Launch::run so that
// launches are counted (bench.py's gpu_launches) and, under option "profile", bracketed by CUDA
// events recorded on the launching stream (bench.py's roofline.achieved).
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <unordered_map>
#include <vector>

namespace ofb {

struct KernelStat { std::string name; uint64_t launches = 0; double total_ms = 0; };

struct Profiler {
    bool timing = false;
    std::vector<KernelStat> stats;
    std::unordered_map<std::string, int> ids;
    struct Pending { int id; cudaEvent_t a, b; };
    std::vector<Pending> pending;
    std::vector<cudaEvent_t> pool;

    int id_of(const char* name)
    {
        auto it = ids.find(name);
        if (it != ids.end()) return it->second;
        int id = (int)stats.size();
        stats.push_back(KernelStat{name, 0, 0});
        ids.emplace(name, id);
        return id;
    }
    cudaEvent_t get_event()
    {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e; cudaEventCreate(&e); return e;
    }
    // Waits for all bracketed launches and folds their durations into stats.
    void collect()
    {
        for (auto& p : pending) {
            cudaEventSynchronize(p.b);
            float ms = 0;
            if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) stats[p.id].total_ms += ms;
            pool.push_back(p.a); pool.push_back(p.b);
        }
        pending.clear();
    }
    void reset() { collect(); for (auto& s : stats) { s.launches = 0; s.total_ms = 0; } }
    ~Profiler()
    {
        for (auto& p : pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
        for (auto e : pool) cudaEventDestroy(e);
    }
};

// Per-context kernel options (ofb_set_option), carried by every Launch: nothing about a launch is process-global,
// so two contexts (two devices, or two host threads) never see each other's settings.
struct KernelOptions {
    int iter_ilp = 1;          // pixels whose UpdateMatrices gathers a k_iter thread keeps in flight
    int iter_prefetch = 1;     // software L2 prefetch one step ahead in k_iter
    int polyexp_tma = 0;       // persistent TMA variant of the scale-0 polynomial expansion
    int generic_polyexp = 0;   // 1 = per-frame part (pyramid + polynomial expansion) through the simple kernels that keep cv2's exact float / double mix
    int pyr_fused = 1;         // scales 1..3 of the pyramid in one pass over the frame (k_pyr_fused)
    int polyexp_fast = 1;      // interior tiles of the polynomial expansion take pe_tile_fast (packed f32x2 vertical pass)
    int polyexp_exact = 1;     // 1 = the horizontal pass of the polynomial expansion in cv2's exact float / double mix (bit-identical to the oracle)
    int exact_window_sums = 1; // 1 = box-window sums with cv2's own arithmetic (k_iter64: double running sums of float differences down the
                               // whole column, double solve): matches cv2 where windows are rank-deficient; ~13 % slower per step
};

struct Launch {
    cudaStream_t stream;
    Profiler* prof;
    int device = 0;            // device the stream belongs to (function attributes are per device)
    int sm_count = 148;
    KernelOptions opt{};

    // cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies to the CURRENT device only: remember, per kernel
    // instance (`done` is a static of the calling template), which devices have been configured.
    template <class K> void dyn_smem(K kernel, size_t bytes, unsigned long long& done) const
    {
        const unsigned long long bit = 1ull << (device & 63);
        if (done & bit) return;
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        done |= bit;
    }
    template <class F> void run(const char* name, F&& f)
    {
        int id = prof->id_of(name);
        prof->stats[id].launches++;
        if (prof->timing) {
            cudaEvent_t a = prof->get_event(), b = prof->get_event();
            cudaEventRecord(a, stream);
            f(stream);
            cudaEventRecord(b, stream);
            prof->pending.push_back({id, a, b});
        } else {
            f(stream);
        }
    }
};

}  // namespace ofb
