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

struct Launch {
    cudaStream_t stream;
    Profiler* prof;
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
