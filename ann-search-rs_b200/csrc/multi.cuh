// multi.cuh -- multi-device handles: one index sharded over several GPUs of one box, driven from ONE process.
//
//   annb_flat_create_multi / annb_ivf_create_multi build one ordinary single-device shard per device --
// contiguous row ranges (flat) or contiguous list ranges balanced by vector count (IVF; centroid table and CSR offsets
// replicated, so every shard derives the same probe lists) -- and return a front handle.  The ordinary entry points
// (annb_flat_search, annb_ivf_search, annb_flat_search_self, the _dev variants with buffers on the first device) accept
// that handle, so the crate's query_*_index_gpu wrappers scale over the box without any change on the Rust side.
//
// One search batch:
//   1. a resident worker thread per device copies the batch's queries straight from the caller's buffer to its device
//      (pinned host memory: one DMA per device, in parallel over each device's own PCIe link; device memory: a peer copy),
//   2. IVF: every device ranks the centroids for ITS slice of the batch only and stores the slice's probe lists into every
//      peer's probe table over NVLink (cudaMemcpyPeerAsync); a host barrier separates the exchange from the scan,
//   3. every device searches its shard for the whole batch (the single-device pipeline, certificate fallback included)
//      and stores its interleaved [ids | distances] block into its slot of the gather buffer on the first device
//      (peer store over NVLink, enqueued on the producing stream right behind the kernel that wrote the block),
//   4. the first device merges the slots under (distance, shard, position) -- the unsharded order -- and the result
//      leaves with one device -> host copy.
// No host bounce, no NCCL: inside one process peer copies are the shortest path over NVSwitch.  Multi-process
// deployments (one process per GPU) use the same per-shard entry points with ncclAllGather, see python/annb200/distributed.py.
#pragma once
#include <condition_variable>
#include <functional>
#include <memory>
#include <thread>

namespace annb {

struct ShardWorker {
    int device = 0;
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    std::function<int()> job;
    bool has_job = false, done = true, quit = false;
    int rc = ANNB_OK;
    std::string err;

    void start() {
        th = std::thread([this] {
            cudaSetDevice(device);
            std::unique_lock<std::mutex> lk(m);
            for (;;) {
                cv.wait(lk, [this] { return has_job || quit; });
                if (quit) return;
                has_job = false;
                lk.unlock();
                int r = ANNB_ERR_CUDA;
                try { r = job(); } catch (const std::exception& e) { g_last_error = std::string("worker exception: ") + e.what(); }
                lk.lock();
                rc = r;
                err = r == ANNB_OK ? std::string() : g_last_error;
                done = true;
                cv.notify_all();
            }
        });
    }
    void post(std::function<int()> f) {
        std::lock_guard<std::mutex> lk(m);
        job = std::move(f);
        has_job = true;
        done = false;
        cv.notify_all();
    }
    int wait() {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [this] { return done; });
        return rc;
    }
    void stop() {
        {
            std::lock_guard<std::mutex> lk(m);
            quit = true;
            cv.notify_all();
        }
        if (th.joinable()) th.join();
    }
};

}  // namespace annb

struct annb_multi {
    std::vector<annb_index*> shards;               // ascending row / list ranges, shards[i] lives on devices[i]
    std::vector<int> devices;
    std::vector<uint64_t> row0;                    // flat: first row of every shard (+ n at the end)
    std::vector<std::unique_ptr<annb::ShardWorker>> workers;
    // per-shard device buffers (grow-only; touched by the shard's worker only)
    std::vector<annb::DevBuf> d_q, d_res, d_probes, d_nprobes, d_bound, d_mdist;
    // first device: gather slots, merged result
    annb::DevBuf g_parts, g_ids, g_dist, g_cnt;
    cudaStream_t root_stream = nullptr;
    std::mutex mu;
};

namespace annb {

static int multi_run_all(annb_multi* m, const std::function<int(size_t)>& f) {
    for (size_t i = 0; i < m->workers.size(); i++) m->workers[i]->post([&f, i] { return f(i); });
    int rc = ANNB_OK;
    std::string err;
    for (size_t i = 0; i < m->workers.size(); i++) {
        const int r = m->workers[i]->wait();
        if (r != ANNB_OK && rc == ANNB_OK) { rc = r; err = "device " + std::to_string(m->devices[i]) + ": " + m->workers[i]->err; }
    }
    if (rc != ANNB_OK) set_last_error(err);
    return rc;
}

static void multi_enable_peers(const std::vector<int>& devices) {
    int prev = 0;
    cudaGetDevice(&prev);
    for (int a : devices) {
        cudaSetDevice(a);
        for (int b : devices) {
            if (a == b) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, a, b) == cudaSuccess && can) cudaDeviceEnablePeerAccess(b, 0);   // "already enabled" is fine
            (void)cudaGetLastError();
        }
    }
    cudaSetDevice(prev);
}

static int multi_check_devices(const int* devices, int n_devices) {
    if (!devices || n_devices <= 0) return fail(ANNB_ERR_INVALID_ARGUMENT, "empty device list");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { (void)cudaGetLastError(); return fail(ANNB_ERR_CUDA, "no usable CUDA device (libannb200 has no CPU fallback)"); }
    for (int i = 0; i < n_devices; i++) {
        // (an ordinal may appear more than once: several shards on one GPU -- how the tests exercise sharding on a one-GPU box)
        if (devices[i] < 0 || devices[i] >= count) return fail(ANNB_ERR_INVALID_ARGUMENT, "device " + std::to_string(devices[i]) + " does not exist");
    }
    return ANNB_OK;
}

static annb_index* multi_front(annb_multi* m, bool ivf) {
    annb_index* f = new annb_index();
    const annb_index* s0 = m->shards[0];
    f->multi = m;
    f->device = m->devices[0];
    f->dtype = s0->dtype; f->metric = s0->metric; f->is_ivf = ivf; f->dim = s0->dim; f->row_bytes = s0->row_bytes;
    f->nlist = s0->nlist; f->list_begin = 0; f->list_end = s0->nlist;
    f->n_total = s0->n_total;
    f->n = 0;
    for (annb_index* s : m->shards) { f->n += s->n; f->device_bytes += s->device_bytes; }
    if (!ivf) f->n_total = f->n;
    return f;
}

static void multi_start_workers(annb_multi* m) {
    const size_t nd = m->devices.size();
    m->d_q.resize(nd); m->d_res.resize(nd); m->d_probes.resize(nd); m->d_nprobes.resize(nd); m->d_bound.resize(nd); m->d_mdist.resize(nd);
    for (size_t i = 0; i < nd; i++) {
        m->workers.emplace_back(new ShardWorker());
        m->workers.back()->device = m->devices[i];
        m->workers.back()->start();
    }
}

static void multi_destroy(annb_multi* m) {
    if (!m) return;
    if (!m->workers.empty()) {
        multi_run_all(m, [m](size_t i) {
            m->d_q[i].release(); m->d_res[i].release(); m->d_probes[i].release(); m->d_nprobes[i].release(); m->d_bound[i].release(); m->d_mdist[i].release();
            return ANNB_OK;
        });
    }
    for (auto& w : m->workers) w->stop();
    for (annb_index* s : m->shards) annb_destroy(s);
    int prev = -1;
    cudaGetDevice(&prev);
    if (!m->devices.empty()) {
        cudaSetDevice(m->devices[0]);
        m->g_parts.release(); m->g_ids.release(); m->g_dist.release(); m->g_cnt.release();
        if (m->root_stream) cudaStreamDestroy(m->root_stream);
    }
    if (prev >= 0) cudaSetDevice(prev);
    (void)cudaGetLastError();
    delete m;
}

// Per-dimension max |x| of one shard's rows (after normalisation for cosine): the shards of an SQ8 index share one
// codebook trained on all rows (ExhaustiveSq8Index::new, src/quantised/exhaustive_sq8.rs:104-151).
static int multi_sq8_absmax(const float* data, uint64_t n, uint32_t dim, bool cosine, std::vector<uint32_t>* out) {
    const uint32_t ld = round_up(dim * 4u, 16u) / 4u;
    DevBuf rows, mx;
    struct Rel { DevBuf& a; DevBuf& b; ~Rel() { a.release(); b.release(); } } rel{rows, mx};
    ANNB_TRY(rows.ensure(n * ld * 4ull));
    ANNB_TRY(mx.ensure(dim * 4ull));
    if (ld != dim) ANNB_CUDA_CHECK(cudaMemset(rows.p, 0, n * ld * 4ull));
    ANNB_CUDA_CHECK(cudaMemcpy2D(rows.p, ld * 4ull, data, dim * 4ull, dim * 4ull, n, cudaMemcpyDefault));
    ANNB_CUDA_CHECK(cudaMemset(mx.p, 0, dim * 4ull));
    if (cosine) normalise_rows_f32_kernel<<<grid_for(n, 128, 1u << 30), 128>>>(rows.as<float>(), ld, dim, n);
    sq8_absmax_kernel<<<dim3(static_cast<uint32_t>(std::min<uint64_t>(n, 2048)), ceil_div(dim, 128u)), 128>>>(rows.as<float>(), ld, dim, n, mx.as<uint32_t>());
    ANNB_CUDA_CHECK(cudaGetLastError());
    out->resize(dim);
    ANNB_CUDA_CHECK(cudaMemcpy(out->data(), mx.p, dim * 4ull, cudaMemcpyDeviceToHost));
    return ANNB_OK;
}

static int multi_flat_create(annb_index** out, const float* data, uint64_t n, uint32_t dim, int dtype, int metric, const int* devices, int n_devices) {
    if (!out) return fail(ANNB_ERR_INVALID_ARGUMENT, "null out");
    *out = nullptr;
    ANNB_TRY(multi_check_devices(devices, n_devices));
    if (!data || n == 0 || dim == 0) return fail(ANNB_ERR_INVALID_ARGUMENT, "empty data");
    if (n < static_cast<uint64_t>(n_devices)) return fail(ANNB_ERR_INVALID_ARGUMENT, "fewer rows than devices");
    annb_multi* m = new annb_multi();
    struct Cleanup { annb_multi*& p; bool armed = true; ~Cleanup() { if (armed) multi_destroy(p); } } cleanup{m};
    m->devices.assign(devices, devices + n_devices);
    m->shards.assign(n_devices, nullptr);
    for (int i = 0; i <= n_devices; i++) m->row0.push_back((static_cast<uint64_t>(i) * n) / n_devices);
    multi_enable_peers(m->devices);
    multi_start_workers(m);
    std::vector<float> scales;
    if (dtype == ANNB_SQ8) {
        std::vector<std::vector<uint32_t>> mx(n_devices);
        ANNB_TRY(multi_run_all(m, [&](size_t i) { return multi_sq8_absmax(data + m->row0[i] * dim, m->row0[i + 1] - m->row0[i], dim, metric == ANNB_COSINE, &mx[i]); }));
        scales.resize(dim);
        for (uint32_t d = 0; d < dim; d++) {
            uint32_t b = 0;
            for (int i = 0; i < n_devices; i++) b = std::max(b, mx[i][d]);
            float f;
            std::memcpy(&f, &b, 4);
            scales[d] = f <= 0.0f ? 1.0f : f / 128.0f;   // ScalarQuantiser::train (src/quantised/quantisers.rs:123-146)
        }
    }
    ANNB_TRY(multi_run_all(m, [&](size_t i) {
        return annb_flat_create(&m->shards[i], data + m->row0[i] * dim, m->row0[i + 1] - m->row0[i], dim, dtype, metric,
                                scales.empty() ? nullptr : scales.data(), m->row0[i], m->devices[i]);
    }));
    {
        DeviceGuard g(m->devices[0]);
        ANNB_CUDA_CHECK(cudaStreamCreateWithFlags(&m->root_stream, cudaStreamNonBlocking));
    }
    *out = multi_front(m, false);
    cleanup.armed = false;
    return ANNB_OK;
}

// Contiguous list ranges with (nearly) equal vector counts: every shard's slab stays one contiguous array.
static std::vector<uint32_t> multi_list_bounds(const std::vector<uint64_t>& off, int parts) {
    const uint32_t nlist = static_cast<uint32_t>(off.size() - 1);
    const uint64_t n = off[nlist];
    std::vector<uint32_t> b{0};
    for (int r = 1; r < parts; r++) {
        const double target = static_cast<double>(n) * r / parts;
        uint32_t c = static_cast<uint32_t>(std::lower_bound(off.begin(), off.end(), static_cast<uint64_t>(std::ceil(target))) - off.begin());
        c = std::min(std::max(c, b.back()), nlist);
        b.push_back(c);
    }
    b.push_back(nlist);
    return b;
}

static int multi_ivf_create(annb_index** out, const void* vectors, const void* norms, const float* centroids, const float* centroid_norms,
                            const uint64_t* offsets, const uint64_t* original_ids, uint64_t n, uint32_t dim, uint32_t nlist, int dtype, int metric,
                            const float* sq8_scales, const int* devices, int n_devices) {
    if (!out) return fail(ANNB_ERR_INVALID_ARGUMENT, "null out");
    *out = nullptr;
    ANNB_TRY(multi_check_devices(devices, n_devices));
    if (!vectors || !centroids || !offsets || !original_ids || n == 0 || dim == 0 || nlist == 0) return fail(ANNB_ERR_INVALID_ARGUMENT, "empty index contents");
    if (dtype < ANNB_F32 || dtype > ANNB_SQ8) return fail(ANNB_ERR_INVALID_ARGUMENT, "unknown dtype");
    std::vector<uint64_t> off(nlist + 1);
    {
        cudaError_t e = cudaMemcpy(off.data(), offsets, (nlist + 1) * 8ull, cudaMemcpyDefault);
        if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(ANNB_ERR_CUDA, std::string("offsets copy: ") + cudaGetErrorString(e)); }
    }
    for (uint32_t c = 0; c < nlist; c++)
        if (off[c] > off[c + 1]) return fail(ANNB_ERR_INVALID_ARGUMENT, "offsets must be non-decreasing");
    if (off[nlist] != n) return fail(ANNB_ERR_INVALID_ARGUMENT, "offsets[nlist] must equal n");
    annb_multi* m = new annb_multi();
    struct Cleanup { annb_multi*& p; bool armed = true; ~Cleanup() { if (armed) multi_destroy(p); } } cleanup{m};
    m->devices.assign(devices, devices + n_devices);
    m->shards.assign(n_devices, nullptr);
    const std::vector<uint32_t> lb = multi_list_bounds(off, n_devices);
    multi_enable_peers(m->devices);
    multi_start_workers(m);
    const uint64_t row = static_cast<uint64_t>(dim) * elem_bytes(dtype);
    ANNB_TRY(multi_run_all(m, [&](size_t i) {
        const uint64_t r0 = off[lb[i]];
        const void* nrm = norms ? static_cast<const void*>(static_cast<const uint8_t*>(norms) + r0 * 4) : nullptr;
        return annb_ivf_create(&m->shards[i], static_cast<const uint8_t*>(vectors) + r0 * row, nrm, centroids, centroid_norms, off.data(), original_ids + r0, n,
                               dim, nlist, dtype, metric, sq8_scales, lb[i], lb[i + 1], m->devices[i]);
    }));
    {
        DeviceGuard g(m->devices[0]);
        ANNB_CUDA_CHECK(cudaStreamCreateWithFlags(&m->root_stream, cudaStreamNonBlocking));
    }
    *out = multi_front(m, true);
    cleanup.armed = false;
    return ANNB_OK;
}

static inline uint64_t multi_res_stride(uint64_t nb, uint32_t k) { return round_up<uint64_t>(nb * k * 12ull, 256); }

// Probe-list pitch of the exchange: room for the probe expansion of select_probed_clusters.
static inline uint32_t multi_probe_pitch(uint32_t np) { return round_up(np + 32u, 32u); }

// mode 0: external f32 queries [nq][dim] (host or device memory);  mode 1: flat self queries, rows [pos_begin, pos_begin + nq)
// out_* : host memory, or device memory of the first device.
static int multi_search(annb_index* front, bool ivf, int mode, const float* queries, uint64_t pos_begin, uint64_t nq, uint32_t k, uint32_t nprobe,
                        uint64_t* out_ids, float* out_dist, uint32_t* out_counts) {
    annb_multi* m = front->multi;
    if (ivf != front->is_ivf) return fail(ANNB_ERR_INVALID_ARGUMENT, ivf ? "not an IVF index" : "not a flat index");
    if (!out_ids || (mode == 0 && !queries && nq)) return fail(ANNB_ERR_INVALID_ARGUMENT, "null buffer");
    if (k == 0) return fail(ANNB_ERR_INVALID_ARGUMENT, "k must be >= 1");
    if (mode == 1 && ivf) return fail(ANNB_ERR_UNSUPPORTED, "self search needs an unsharded IVF index");
    std::lock_guard<std::mutex> lock(m->mu);
    const size_t nd = m->shards.size();
    const uint32_t dim = front->dim;
    const int root = m->devices[0];
    ANNB_DEVICE(root);
    uint32_t np = nprobe ? nprobe : std::max<uint32_t>(1, static_cast<uint32_t>(std::sqrt(static_cast<double>(front->nlist))));
    np = std::min(np, std::max(front->nlist, 1u));
    for (uint64_t b0 = 0; b0 < nq; b0 += QUERY_BATCH) {
        const uint64_t nb = std::min<uint64_t>(QUERY_BATCH, nq - b0);
        const uint64_t stride = multi_res_stride(nb, k);
        ANNB_TRY(m->g_parts.ensure(stride * nd));
        ANNB_TRY(m->g_ids.ensure(nb * k * 8ull));
        ANNB_TRY(m->g_dist.ensure(nb * k * 4ull));
        ANNB_TRY(m->g_cnt.ensure(nb * 4ull));
        uint8_t* gather = m->g_parts.as<uint8_t>();
        // ---- queries onto every device ----
        std::vector<PreparedQueries> pqs(nd);
        ANNB_TRY(multi_run_all(m, [&](size_t i) -> int {
            annb_index* ix = m->shards[i];
            std::lock_guard<std::mutex> sl(ix->mu);
            cudaStream_t s = ix->stream;
            ANNB_TRY(order_after_previous(ix, s));
            if (mode == 0) {
                ANNB_TRY(m->d_q[i].ensure(nb * dim * 4ull));
                ANNB_CUDA_CHECK(cudaMemcpyAsync(m->d_q[i].p, queries + b0 * dim, nb * dim * 4ull, cudaMemcpyDefault, s));
                ANNB_TRY(prepare_external(ix, m->d_q[i].as<float>(), nb, &pqs[i], s));
            } else {
                // the query rows, in the index dtype, come from the shards that own them (peer copies)
                ANNB_TRY(m->d_q[i].ensure(nb * static_cast<uint64_t>(ix->row_bytes)));
                const uint64_t a = pos_begin + b0, b = a + nb;
                for (size_t o = 0; o < nd; o++) {
                    const uint64_t lo = std::max(a, m->row0[o]), hi = std::min(b, m->row0[o + 1]);
                    if (lo >= hi) continue;
                    ANNB_CUDA_CHECK(cudaMemcpyPeerAsync(m->d_q[i].as<uint8_t>() + (lo - a) * ix->row_bytes, m->devices[i],
                                                        m->shards[o]->d_rows + (lo - m->row0[o]) * ix->row_bytes, m->devices[o], (hi - lo) * ix->row_bytes, s));
                }
                PreparedQueries& pq = pqs[i];
                pq.scan = m->d_q[i].as<uint8_t>(); pq.scan_bytes = ix->row_bytes;
                pq.qt = ix->dtype == ANNB_F32 ? QT_F32 : (ix->dtype == ANNB_BF16 ? QT_BF16 : QT_I8);
                pq.bf16_self = ix->dtype == ANNB_BF16;
                pq.route = nullptr;
            }
            return ANNB_OK;
        }));
        // ---- IVF: every device ranks the centroids for its slice of the batch and stores the probe lists on every peer ----
        uint32_t pitch = 0;
        if (ivf) {
            pitch = multi_probe_pitch(np);
            const uint64_t per = ceil_div<uint64_t>(nb, nd);
            ANNB_TRY(multi_run_all(m, [&](size_t i) -> int {
                ANNB_TRY(m->d_probes[i].ensure(nb * static_cast<uint64_t>(pitch) * 4));
                ANNB_TRY(m->d_nprobes[i].ensure(nb * 4ull));
                return ANNB_OK;
            }));
            ANNB_TRY(multi_run_all(m, [&](size_t i) -> int {
                annb_index* ix = m->shards[i];
                std::lock_guard<std::mutex> sl(ix->mu);
                cudaStream_t s = ix->stream;
                const uint64_t lo = std::min<uint64_t>(nb, i * per), hi = std::min<uint64_t>(nb, (i + 1) * per);
                if (hi <= lo) return ANNB_OK;
                PreparedQueries sl_pq = pqs[i];
                sl_pq.scan += lo * sl_pq.scan_bytes;
                sl_pq.route += lo * sl_pq.route_ld;
                uint32_t* my_p = m->d_probes[i].as<uint32_t>() + lo * pitch;
                uint32_t* my_n = m->d_nprobes[i].as<uint32_t>() + lo;
                ANNB_TRY(route_batch(ix, sl_pq, hi - lo, k, nprobe, RouteOut{my_p, my_n, pitch}, s));
                for (size_t o = 0; o < nd; o++) {
                    if (o == i) continue;
                    ANNB_CUDA_CHECK(cudaMemcpyPeerAsync(m->d_probes[o].as<uint32_t>() + lo * pitch, m->devices[o], my_p, m->devices[i], (hi - lo) * pitch * 4ull, s));
                    ANNB_CUDA_CHECK(cudaMemcpyPeerAsync(m->d_nprobes[o].as<uint32_t>() + lo, m->devices[o], my_n, m->devices[i], (hi - lo) * 4ull, s));
                }
                ANNB_CUDA_CHECK(cudaStreamSynchronize(s));
                return ANNB_OK;
            }));
        }
        // ---- every device searches its shard for the whole batch; its result block lands in its gather slot ----
        // External queries run in shard mode: no shard certifies its own k-th neighbour, every shard reports a bound that is
        // tested against the merged result below.  Self queries (index dtype) keep the local certificate.
        const bool shard_mode = mode == 0;
        cudaStream_t rs = m->root_stream;
        auto copy_block = [&](size_t i, cudaStream_t s) -> int {
            ANNB_CUDA_CHECK(cudaMemcpyPeerAsync(gather + i * stride, root, m->d_res[i].p, m->devices[i], nb * k * 12ull, s));
            return ANNB_OK;
        };
        ANNB_TRY(multi_run_all(m, [&](size_t i) -> int {
            annb_index* ix = m->shards[i];
            cudaStream_t s = ix->stream;
            ANNB_TRY(m->d_res[i].ensure(stride));
            uint64_t* r_ids = m->d_res[i].as<uint64_t>();
            float* r_dist = reinterpret_cast<float*>(m->d_res[i].as<uint8_t>() + nb * k * 8ull);
            if (ix->n == 0) {   // a shard without vectors (more devices than lists): an all-sentinel block
                ANNB_CUDA_CHECK(cudaMemsetAsync(m->d_res[i].p, 0xFF, nb * k * 8ull, s));
                ANNB_TRY(copy_block(i, s));
                ANNB_CUDA_CHECK(cudaStreamSynchronize(s));
                return ANNB_OK;
            }
            if (shard_mode) {
                ANNB_TRY(m->d_bound[i].ensure(nb * 4ull));
                ANNB_TRY(m->d_mdist[i].ensure(nb * k * 4ull));
                if (ivf) ANNB_TRY(annb_ivf_search_probes_shard_dev(ix, m->d_q[i].as<float>(), nb, dim, k, nprobe, m->d_probes[i].as<uint32_t>(),
                                                                   m->d_nprobes[i].as<uint32_t>(), pitch, r_ids, r_dist, m->d_bound[i].as<float>(), s));
                else ANNB_TRY(annb_flat_search_shard_dev(ix, m->d_q[i].as<float>(), nb, dim, k, r_ids, r_dist, m->d_bound[i].as<float>(), s));
                ANNB_TRY(copy_block(i, s));
                ANNB_CUDA_CHECK(cudaStreamSynchronize(s));
                return ANNB_OK;
            }
            std::lock_guard<std::mutex> sl(ix->mu);
            auto copy_out = [&]() -> int { return copy_block(i, s); };
            ANNB_TRY(run_batch(ix, false, pqs[i], nb, k, 0, r_ids, r_dist, nullptr, s, true, copy_out));
            return mark_call_done(ix, s);
        }));
        ANNB_TRY(annb_merge_shards_dev(gather, stride, nb * k * 8ull, static_cast<uint32_t>(nd), nb, k, m->g_ids.as<uint64_t>(), m->g_dist.as<float>(),
                                       m->g_cnt.as<uint32_t>(), rs));
        if (shard_mode) {
            // the merged k-th distances go back to every device; a shard whose bound does not clear them recomputes those queries
            // exactly and re-sends its block, and the merge is repeated (rare)
            ANNB_CUDA_CHECK(cudaStreamSynchronize(rs));
            std::vector<uint32_t> counts(nd, 0);
            ANNB_TRY(multi_run_all(m, [&](size_t i) -> int {
                annb_index* ix = m->shards[i];
                if (ix->n == 0) return ANNB_OK;
                cudaStream_t s = ix->stream;
                ANNB_CUDA_CHECK(cudaMemcpyPeerAsync(m->d_mdist[i].p, m->devices[i], m->g_dist.p, root, nb * k * 4ull, s));
                ANNB_TRY(annb_shard_check_dev(ix, m->d_bound[i].as<float>(), m->d_mdist[i].as<float>(), nb, k, &counts[i], s));
                if (counts[i] == 0) return ANNB_OK;
                uint64_t* r_ids = m->d_res[i].as<uint64_t>();
                float* r_dist = reinterpret_cast<float*>(m->d_res[i].as<uint8_t>() + nb * k * 8ull);
                ANNB_TRY(annb_shard_refine_dev(ix, m->d_q[i].as<float>(), nb, dim, k, nprobe, ivf ? m->d_probes[i].as<uint32_t>() : nullptr,
                                               ivf ? m->d_nprobes[i].as<uint32_t>() : nullptr, pitch, r_ids, r_dist, s));
                ANNB_TRY(copy_block(i, s));
                ANNB_CUDA_CHECK(cudaStreamSynchronize(s));
                return ANNB_OK;
            }));
            bool any = false;
            for (uint32_t c : counts) any = any || c != 0;
            if (any)
                ANNB_TRY(annb_merge_shards_dev(gather, stride, nb * k * 8ull, static_cast<uint32_t>(nd), nb, k, m->g_ids.as<uint64_t>(),
                                               m->g_dist.as<float>(), m->g_cnt.as<uint32_t>(), rs));
        }
        // ---- one copy out of the merged rows ----
        ANNB_CUDA_CHECK(cudaMemcpyAsync(out_ids + b0 * k, m->g_ids.p, nb * k * 8ull, cudaMemcpyDefault, rs));
        if (out_dist) ANNB_CUDA_CHECK(cudaMemcpyAsync(out_dist + b0 * k, m->g_dist.p, nb * k * 4ull, cudaMemcpyDefault, rs));
        if (out_counts) ANNB_CUDA_CHECK(cudaMemcpyAsync(out_counts + b0, m->g_cnt.p, nb * 4ull, cudaMemcpyDefault, rs));
        ANNB_CUDA_CHECK(cudaStreamSynchronize(rs));
    }
    return ANNB_OK;
}

}  // namespace annb
