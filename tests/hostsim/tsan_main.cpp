// TEST INFRASTRUCTURE ONLY. Driver for a ThreadSanitizer build of the host emulation (tests/test_hostsim_tsan.py): the lane groups of
// roc_group.cuh as free-running host threads, skewed id sets (overflowing buckets, spill list, brute-force fallback).
#include <cstdint>
#include <cstdio>
#include <vector>
#include <random>
#include <algorithm>
extern "C" {
int64_t sim_group_encode(int G, uint32_t n, const uint64_t* ids, int prec, uint64_t* head_out, uint32_t* words_out,
                         uint32_t cap, uint32_t* order_out, uint32_t* status_out);
void sim_group_decode(int G, uint64_t head, const uint32_t* words, uint32_t nwords, uint32_t n, int prec, uint32_t lo,
                      uint32_t hi, int64_t* out, uint32_t* status_out, uint32_t force_degenerate);
int64_t sim_warp_encode(uint32_t n, const uint64_t* ids, int prec, uint64_t* head_out, uint32_t* words_out, uint32_t cap,
                        uint32_t* order_out, uint32_t* status_out);
void sim_warp_decode(uint64_t head, const uint32_t* words, uint32_t nwords, uint32_t n, int prec, int row, int64_t* out,
                     uint32_t* status_out);
}
int main() {
    std::mt19937 g(7);
    for (int G : {2, 4, 8}) {
        for (uint32_t n : {1u, 50u, 900u, 3000u}) {
            std::vector<uint64_t> ids;
            // skewed: half the ids in a narrow cluster -> overflowing buckets, spill list
            while (ids.size() < n) ids.push_back((g() % 2) ? g() % 5000 : g() % (1u << 20));
            std::sort(ids.begin(), ids.end()); ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
            uint32_t m = (uint32_t)ids.size();
            std::vector<uint32_t> words(m + 4), order(m + 1); uint64_t head = 0; uint32_t st = 0;
            int64_t nw = sim_group_encode(G, m, ids.data(), 20, &head, words.data(), m + 4, order.data(), &st);
            std::vector<int64_t> out(m + 1); uint32_t st2 = 0;
            sim_group_decode(G, head, words.data(), (uint32_t)nw, m, 20, (uint32_t)ids.front(), (uint32_t)ids.back(), out.data(), &st2, 0);
            std::vector<uint64_t> back(out.begin(), out.begin() + m); std::sort(back.begin(), back.end());
            printf("G=%d n=%u words=%lld status=%u/%u roundtrip=%s\n", G, m, (long long)nw, st, st2, back == ids ? "ok" : "MISMATCH");
        }
    }
    // one unit per warp (roc_small.cuh): shared-memory ids written by lane 0, read by all lanes a rendezvous later; the
    // encoder's stream words stored by lane 0, re-read by every lane's refill
    for (uint32_t n : {40u, 150u}) {
        std::vector<uint64_t> ids;
        while (ids.size() < n) ids.push_back(g() % (1u << 20));
        std::sort(ids.begin(), ids.end()); ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
        uint32_t m = (uint32_t)ids.size();
        std::vector<uint32_t> words(m + 4), order(m + 1); uint64_t head = 0; uint32_t st = 0;
        int64_t nw = sim_warp_encode(m, ids.data(), 20, &head, words.data(), m + 4, order.data(), &st);
        for (int row = 0; row < (m <= 64 ? 2 : 1); row++) {
            std::vector<int64_t> out(m + 1); uint32_t st2 = 0;
            sim_warp_decode(head, words.data(), (uint32_t)nw, m, 20, row, out.data(), &st2);
            std::vector<uint64_t> back(out.begin(), out.begin() + m); std::sort(back.begin(), back.end());
            printf("W=32 row=%d n=%u words=%lld status=%u/%u roundtrip=%s\n", row, m, (long long)nw, st, st2, back == ids ? "ok" : "MISMATCH");
        }
    }
}
