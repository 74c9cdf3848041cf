// TEST INFRASTRUCTURE ONLY. Host emulation of a sub-warp lane group for csrc/roc_group.cuh: every lane is a host
// thread, a collective (ballot / shuffle) is a rendezvous of the group's G threads. Nothing in the product uses it.
#pragma once
#include <atomic>
#include <cstdint>
#include <functional>
#include <thread>
#include <vector>

struct GrpShared {
    int G;
    std::atomic<int> count{0};
    std::atomic<int> sense{0};
    volatile uint32_t slot[32];
};

struct HostGrp {
    uint32_t sub;
    GrpShared* sh;
    mutable int local = 0;
    void barrier() const {
        local ^= 1;
        if (sh->count.fetch_add(1, std::memory_order_acq_rel) + 1 == sh->G) {
            sh->count.store(0, std::memory_order_relaxed);
            sh->sense.store(local, std::memory_order_release);
        } else {
            int spins = 0;
            while (sh->sense.load(std::memory_order_acquire) != local)
                if (++spins > 2000) std::this_thread::yield();
        }
    }
    uint32_t ballot(bool p) const {
        sh->slot[sub] = p ? 1u : 0u;
        barrier();
        uint32_t m = 0;
        for (int j = 0; j < sh->G; j++) m |= sh->slot[j] << j;
        barrier();
        return m;
    }
    uint32_t shfl(uint32_t v, uint32_t src) const {
        sh->slot[sub] = v;
        barrier();
        uint32_t r = sh->slot[src % (uint32_t)sh->G];
        barrier();
        return r;
    }
    uint32_t shfl_xor(uint32_t v, uint32_t m) const { return shfl(v, sub ^ m); }
    void sync() const { barrier(); }
    bool warp_any(bool p) const { return ballot(p) != 0u; }  // the emulated warp holds one group
    void host_sync() const { barrier(); }
};

// run body(lane) on G threads
inline void run_group(int G, const std::function<void(const HostGrp&)>& body) {
    GrpShared sh;
    sh.G = G;
    std::vector<std::thread> th;
    for (int j = 1; j < G; j++) th.emplace_back([&, j] { HostGrp g{(uint32_t)j, &sh}; body(g); });
    HostGrp g0{0u, &sh};
    body(g0);
    for (auto& t : th) t.join();
}
