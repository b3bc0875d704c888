// TEST INFRASTRUCTURE ONLY — a minimal parallel-for on std::thread (this image's gcc ships without libgomp).
// Used only by the functions that say so; every use hands out INDEPENDENT index ranges (no reduction across threads), so results do
// not depend on the thread count.  SRK_ORACLE_THREADS overrides std::thread::hardware_concurrency().
#pragma once
#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <thread>
#include <vector>

namespace srk_oracle {

inline int OracleThreads() {
    if (const char* e = std::getenv("SRK_ORACLE_THREADS")) { int v = std::atoi(e); if (v > 0) return v; }
    unsigned h = std::thread::hardware_concurrency();
    return h == 0 ? 1 : (int)h;
}

// fn(i) for i in [begin, end), chunks of `chunk` indices handed out dynamically
template <class Fn>
inline void ParallelFor(int64_t begin, int64_t end, int64_t chunk, Fn&& fn) {
    const int64_t count = end - begin;
    if (count <= 0) return;
    int nt = OracleThreads();
    if ((int64_t)nt > (count + chunk - 1) / chunk) nt = (int)((count + chunk - 1) / chunk);
    if (nt <= 1) { for (int64_t i = begin; i < end; ++i) fn(i); return; }
    std::atomic<int64_t> next(begin);
    auto work = [&]() {
        for (;;) {
            int64_t s = next.fetch_add(chunk);
            if (s >= end) return;
            int64_t e = s + chunk < end ? s + chunk : end;
            for (int64_t i = s; i < e; ++i) fn(i);
        }
    };
    std::vector<std::thread> th;
    th.reserve((size_t)nt - 1);
    for (int t = 1; t < nt; ++t) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
}

}  // namespace srk_oracle
