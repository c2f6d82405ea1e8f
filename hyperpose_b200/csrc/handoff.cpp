// handoff.cpp -- registry of published engine output batches (see handoff.h).
#include "handoff.h"

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <unordered_map>

#include "common.h"

namespace hpb {
namespace handoff {

namespace {

std::mutex g_mu; // guards g_map; lock order: Batch::mu before g_mu
std::unordered_map<const float*, std::pair<std::weak_ptr<Batch>, int>> g_map;
std::atomic<long long> g_published{ 0 }, g_hits{ 0 }, g_batch_parses{ 0 }, g_misses{ 0 };
std::atomic<int> g_enabled{ -1 }; // -1: not decided yet (HPB_NO_HANDOFF)

void unregister_locked(Batch& b) // caller holds b.mu
{
    std::lock_guard<std::mutex> lk(g_mu);
    for (const float* h : b.host_a) {
        auto it = g_map.find(h);
        if (it == g_map.end()) continue;
        auto owner = it->second.first.lock();
        if (!owner || owner.get() == &b) g_map.erase(it);
    }
}

} // namespace

bool enabled()
{
    int v = g_enabled.load();
    if (v < 0) {
        v = std::getenv("HPB_NO_HANDOFF") ? 0 : 1;
        g_enabled.store(v);
    }
    return v != 0;
}
void set_enabled(bool on) { g_enabled.store(on ? 1 : 0); }

bool contents_match(const Batch& b, int frame)
{
    if (!b.pin) return false;
    const float* pa = b.pin + (size_t)frame * b.elems_a;
    const float* pb = b.pin + (size_t)b.N * b.elems_a + (size_t)frame * b.elems_b;
    // bit patterns, not float compares: NaNs must match themselves
    return std::memcmp(b.host_a[frame], pa, b.elems_a * sizeof(float)) == 0 && std::memcmp(b.host_b[frame], pb, b.elems_b * sizeof(float)) == 0;
}

int publish(std::shared_ptr<Batch>* ring, int* ring_pos, int device, cudaStream_t st, const float* d_a, const float* d_b, int N,
            size_t elems_a, size_t elems_b, float* const* host_a, float* const* host_b)
{
    std::shared_ptr<Batch>& slot = ring[*ring_pos];
    *ring_pos = (*ring_pos + 1) % HANDOFF_RING;
    if (!slot) slot = std::make_shared<Batch>();
    Batch& b = *slot;
    std::lock_guard<std::mutex> lk(b.mu); // waits for a parse still reading this slot's snapshot
    if (b.valid) unregister_locked(b);
    b.valid = false;
    b.fail_count = 0;
    b.cache_kind = 0;
    b.device = device;
    const size_t need_a = (size_t)N * elems_a, need_b = (size_t)N * elems_b;
    if (need_a > b.cap_a) {
        if (b.d_a) cudaFree(b.d_a);
        b.d_a = nullptr; b.cap_a = 0;
        HP_CUDA_TRY(cudaMalloc(&b.d_a, need_a * sizeof(float)));
        b.cap_a = need_a;
    }
    if (need_b > b.cap_b) {
        if (b.d_b) cudaFree(b.d_b);
        b.d_b = nullptr; b.cap_b = 0;
        HP_CUDA_TRY(cudaMalloc(&b.d_b, need_b * sizeof(float)));
        b.cap_b = need_b;
    }
    if (need_a + need_b > b.pin_floats) {
        if (b.pin) cudaFreeHost(b.pin);
        b.pin = nullptr; b.pin_floats = 0;
        HP_CUDA_TRY(cudaMallocHost(&b.pin, (need_a + need_b) * sizeof(float)));
        b.pin_floats = need_a + need_b;
    }
    if (!b.ready) HP_CUDA_TRY(cudaEventCreateWithFlags(&b.ready, cudaEventDisableTiming));
    HP_CUDA_TRY(cudaMemcpyAsync(b.pin, d_a, need_a * sizeof(float), cudaMemcpyDeviceToHost, st));
    HP_CUDA_TRY(cudaMemcpyAsync(b.pin + need_a, d_b, need_b * sizeof(float), cudaMemcpyDeviceToHost, st));
    HP_CUDA_TRY(cudaMemcpyAsync(b.d_a, d_a, need_a * sizeof(float), cudaMemcpyDeviceToDevice, st));
    HP_CUDA_TRY(cudaMemcpyAsync(b.d_b, d_b, need_b * sizeof(float), cudaMemcpyDeviceToDevice, st));
    HP_CUDA_TRY(cudaEventRecord(b.ready, st));
    HP_CUDA_TRY(cudaStreamSynchronize(st));
    b.N = N; b.elems_a = elems_a; b.elems_b = elems_b;
    b.host_a.assign(host_a, host_a + N);
    b.host_b.assign(host_b, host_b + N);
    for (int i = 0; i < N; ++i) {   // the caller's per-image buffers (feature_map_t storage) receive the published bytes
        std::memcpy(host_a[i], b.pin + (size_t)i * elems_a, elems_a * sizeof(float));
        std::memcpy(host_b[i], b.pin + need_a + (size_t)i * elems_b, elems_b * sizeof(float));
    }
    b.valid = true;
    {
        std::lock_guard<std::mutex> lg(g_mu);
        for (int i = 0; i < N; ++i) g_map[host_a[i]] = { slot, i };
    }
    g_published.fetch_add(1);
    return HP_OK;
}

void retire_ring(std::shared_ptr<Batch>* ring)
{
    for (int i = 0; i < HANDOFF_RING; ++i) {
        if (!ring[i]) continue;
        Batch& b = *ring[i];
        std::lock_guard<std::mutex> lk(b.mu);
        if (b.valid) unregister_locked(b);
        b.valid = false;
        if (b.d_a) cudaFree(b.d_a);
        if (b.d_b) cudaFree(b.d_b);
        if (b.ready) cudaEventDestroy(b.ready);
        if (b.pin) cudaFreeHost(b.pin);
        b.pin = nullptr; b.pin_floats = 0;
        b.d_a = b.d_b = nullptr;
        b.cap_a = b.cap_b = 0;
        b.ready = nullptr;
    }
    for (int i = 0; i < HANDOFF_RING; ++i) ring[i].reset();
}

Hit lookup(const float* host_a, const float* host_b, size_t elems_a, size_t elems_b)
{
    Hit h;
    if (!enabled()) return h;
    std::lock_guard<std::mutex> lg(g_mu);
    auto it = g_map.find(host_a);
    if (it == g_map.end()) return h;
    auto sp = it->second.first.lock();
    if (!sp) { g_map.erase(it); return h; }
    const int f = it->second.second;
    // geometry and the second pointer are immutable while the entry is registered (publish() unregisters under b.mu first);
    // they are re-checked under b.mu by the caller together with the contents
    if (f >= (int)sp->host_b.size() || sp->host_b[f] != host_b || sp->elems_a != elems_a || sp->elems_b != elems_b) return h;
    h.batch = std::move(sp);
    h.frame = f;
    return h;
}

std::mutex& registry_mutex() { return g_mu; }
int device_of_locked(const float* host_a)
{
    auto it = g_map.find(host_a);
    if (it == g_map.end()) return -1;
    auto sp = it->second.first.lock();
    return sp ? sp->device : -1;   // Batch::device is written under g_mu-free b.mu, but only while the entry is unregistered
}

void count_hit() { g_hits.fetch_add(1); }
void count_batch_parse() { g_batch_parses.fetch_add(1); }
void count_miss() { g_misses.fetch_add(1); }
void stats(long long* published, long long* hits, long long* batch_parses, long long* misses)
{
    if (published) *published = g_published.load();
    if (hits) *hits = g_hits.load();
    if (batch_parses) *batch_parses = g_batch_parses.load();
    if (misses) *misses = g_misses.load();
}

} // namespace handoff
} // namespace hpb

extern "C" {
int hp_handoff_enable(int on)
{
    hpb::handoff::set_enabled(on != 0);
    return HP_OK;
}
int hp_handoff_device_of(const float* host_conf)
{
    if (!host_conf || !hpb::handoff::enabled()) return -1;
    std::lock_guard<std::mutex> lg(hpb::handoff::registry_mutex());
    return hpb::handoff::device_of_locked(host_conf);
}
int hp_handoff_stats(long long* published, long long* hits, long long* batch_parses, long long* misses)
{
    hpb::handoff::stats(published, hits, batch_parses, misses);
    return HP_OK;
}
}
