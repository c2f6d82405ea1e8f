// pool.cpp -- frame-sharded multi-GPU execution inside ONE process: one host thread, one engine, one parser
// and one CUDA stream pair per GPU; weights replicated; no data-path collective.
//
// Partitioning: a call with N_total frames is cut into blocks of B = max_batch frames; block k holds frames
// [k*B, (k+1)*B) and runs on GPU k % n_gpus, i.e. GPU g gets frames [g*B, (g+1)*B) of every super-batch of n_gpus*B
// frames.  Each worker pipelines its blocks two deep through hp_pose_submit_u8_host / hp_pose_collect (H2D of its next
// block under the convs of the current one).  Results are written straight into the caller's arrays at the frames'
// own positions, so they come back in frame order whatever the completion order of the GPUs.
//
// Everything here sits ABOVE the single-GPU C ABI (hp_engine_*, hp_paf_*, hp_pose_*): the pool adds threads and the
// block arithmetic, nothing else.  hp_default_device() is the device selector of the C++ drop-in classes.
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/hyperpose_b200.h"
#include "common.h"

struct hp_pool {
    struct Worker {
        int device = 0;
        hp_engine* engine = nullptr;
        hp_paf* parser = nullptr;
        std::thread th;
        int rc = HP_OK;
        std::string err;
    };
    std::vector<Worker> workers;
    int max_batch = 0, in_w = 0, in_h = 0;
    size_t frame_bytes = 0;
    // one job at a time (hp_pool_run_u8_host is synchronous); generation counter wakes the workers
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    long long generation = 0;
    int pending = 0;
    bool stop = false;
    const uint8_t* frames = nullptr;
    int n_total = 0, cap = 0;
    hp_human* out = nullptr;
    int* n_out = nullptr;
};

namespace {

// blocks g, g + n, g + 2n, ... of the current job, two in flight
void run_blocks(hp_pool* p, int g)
{
    hp_pool::Worker& w = p->workers[g];
    const int n_gpus = (int)p->workers.size(), B = p->max_batch;
    const int n_blocks = (p->n_total + B - 1) / B;
    struct Flight { int ticket, block; };
    Flight fl[2];
    int n_fl = 0;
    auto collect = [&](const Flight& f) {
        const int first = f.block * B;
        const int rc = hp_pose_collect(w.engine, f.ticket, p->out + (size_t)first * p->cap, p->cap, p->n_out + first);
        if (rc != HP_OK && w.rc == HP_OK) { w.rc = rc; w.err = hp_last_error(); }
    };
    for (int k = g; k < n_blocks; k += n_gpus) {
        const int first = k * B, n = std::min(B, p->n_total - first);
        if (n_fl == 2) { collect(fl[0]); fl[0] = fl[1]; n_fl = 1; }
        int ticket = -1;
        const int rc = hp_pose_submit_u8_host(w.engine, w.parser, p->frames + (size_t)first * p->frame_bytes, n, &ticket);
        if (rc != HP_OK) { if (w.rc == HP_OK) { w.rc = rc; w.err = hp_last_error(); } break; }
        fl[n_fl++] = { ticket, k };
    }
    for (int i = 0; i < n_fl; ++i) collect(fl[i]);
}

void worker_main(hp_pool* p, int g)
{
    long long seen = 0;
    for (;;) {
        {
            std::unique_lock<std::mutex> lk(p->mu);
            p->cv_job.wait(lk, [&] { return p->stop || p->generation != seen; });
            if (p->stop) return;
            seen = p->generation;
        }
        run_blocks(p, g);
        {
            std::lock_guard<std::mutex> lk(p->mu);
            if (--p->pending == 0) p->cv_done.notify_all();
        }
    }
}

std::atomic<int> g_rr{ 0 };

} // namespace

extern "C" {

// Device selector of the C++ drop-in (hyperpose::dnn::tensorrt has no device argument, tensorrt.hpp:44-74):
//   HPB_DEVICE=<ordinal>   every engine on that GPU
//   HPB_DEVICE=rr          engine instances take the GPUs round-robin (one engine per stream / thread => one GPU each)
//   unset                  GPU 0, like the reference's cudaSetDevice-less TensorRT code
int hp_default_device(void)
{
    const char* v = std::getenv("HPB_DEVICE");
    if (!v || !*v) return 0;
    const int n = hp_device_count();
    if (n <= 0) return 0;
    if (std::strcmp(v, "rr") == 0 || std::strcmp(v, "round_robin") == 0) return g_rr.fetch_add(1) % n;
    const int d = std::atoi(v);
    return (d >= 0 && d < n) ? d : 0;
}

int hp_pool_create(hp_pool** out, const int* devices, int n_devices, const void* pack, size_t pack_bytes, int in_w, int in_h,
                   int max_batch, double factor, int flip_rgb, float conf_thresh, float paf_thresh)
{
    if (!out || !pack || n_devices <= 0 || n_devices > 64) { hpb::set_error("hp_pool_create: bad argument"); return HP_ERR_ARG; }
    *out = nullptr;
    hp_pool* p = new hp_pool();
    p->max_batch = max_batch; p->in_w = in_w; p->in_h = in_h;
    p->frame_bytes = (size_t)in_h * in_w * 3;
    p->workers.resize(n_devices);
    // engines are built concurrently (weight repacking is host work): one short-lived thread per GPU
    std::vector<std::thread> builders;
    for (int g = 0; g < n_devices; ++g) {
        p->workers[g].device = devices ? devices[g] : g;
        builders.emplace_back([p, g, pack, pack_bytes, in_w, in_h, max_batch, factor, flip_rgb, conf_thresh, paf_thresh] {
            hp_pool::Worker& w = p->workers[g];
            w.rc = hp_engine_create(&w.engine, pack, pack_bytes, in_w, in_h, max_batch, factor, flip_rgb, w.device);
            if (w.rc == HP_OK) w.rc = hp_paf_create(&w.parser, conf_thresh, paf_thresh, -1, -1, w.device);
            if (w.rc != HP_OK) w.err = hp_last_error();
        });
    }
    for (auto& b : builders) b.join();
    for (int g = 0; g < n_devices; ++g)
        if (p->workers[g].rc != HP_OK) {
            const int rc = p->workers[g].rc;
            hpb::set_error("hp_pool_create: GPU %d: %s", p->workers[g].device, p->workers[g].err.c_str());
            for (auto& w : p->workers) { hp_paf_destroy(w.parser); hp_engine_destroy(w.engine); }
            delete p;
            return rc;
        }
    for (int g = 0; g < n_devices; ++g) p->workers[g].th = std::thread(worker_main, p, g);
    *out = p;
    return HP_OK;
}

void hp_pool_destroy(hp_pool* p)
{
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(p->mu);
        p->stop = true;
    }
    p->cv_job.notify_all();
    for (auto& w : p->workers) {
        if (w.th.joinable()) w.th.join();
        hp_paf_destroy(w.parser);
        hp_engine_destroy(w.engine);
    }
    delete p;
}

int hp_pool_size(const hp_pool* p) { return p ? (int)p->workers.size() : 0; }

int hp_pool_set_capacity(hp_pool* p, int max_peaks_per_part, int max_candidates_per_limb, int max_humans)
{
    if (!p) return HP_ERR_ARG;
    for (auto& w : p->workers) {
        const int rc = hp_paf_set_capacity(w.parser, max_peaks_per_part, max_candidates_per_limb, max_humans);
        if (rc) return rc;
    }
    return HP_OK;
}

// benchmark hook: every GPU's engine gets the same override rule -- d_conf[g] / d_paf[g] are DEVICE pointers on GPU g
int hp_pool_set_output_override(hp_pool* p, const float* const* d_conf, const float* const* d_paf)
{
    if (!p) return HP_ERR_ARG;
    for (size_t g = 0; g < p->workers.size(); ++g) {
        const int rc = hp_engine_set_output_override(p->workers[g].engine, d_conf ? d_conf[g] : nullptr, d_paf ? d_paf[g] : nullptr);
        if (rc) return rc;
    }
    return HP_OK;
}

int hp_pool_run_u8_host(hp_pool* p, const uint8_t* frames, int n_total, hp_human* out, int cap, int* n_out)
{
    if (!p || !frames || !out || !n_out || n_total <= 0 || cap <= 0) { hpb::set_error("hp_pool_run_u8_host: bad argument"); return HP_ERR_ARG; }
    {
        std::lock_guard<std::mutex> lk(p->mu);
        p->frames = frames; p->n_total = n_total; p->out = out; p->cap = cap; p->n_out = n_out;
        for (auto& w : p->workers) { w.rc = HP_OK; w.err.clear(); }
        p->pending = (int)p->workers.size();
        ++p->generation;
    }
    p->cv_job.notify_all();
    {
        std::unique_lock<std::mutex> lk(p->mu);
        p->cv_done.wait(lk, [&] { return p->pending == 0; });
    }
    for (auto& w : p->workers)
        if (w.rc != HP_OK) { hpb::set_error("hp_pool_run_u8_host: GPU %d: %s", w.device, w.err.c_str()); return w.rc; }
    return HP_OK;
}

long long hp_pool_launch_count(const hp_pool* p)
{
    long long n = 0;
    if (p) for (auto& w : p->workers) n += hp_engine_launch_count(w.engine) + hp_paf_launch_count(w.parser);
    return n;
}

} // extern "C"
