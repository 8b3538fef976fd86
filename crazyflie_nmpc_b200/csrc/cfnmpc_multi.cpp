// Multi-device batch handle (SURVEY.md 8b surface C / 8e): the batch dimension sharded over several GPUs of one node
// behind ONE C handle -- contiguous slices of B / n_dev instances (the first B mod n_dev devices take one more), one
// cfnmpc_batch + CUDA stream per device, one host thread per device for every call, no exchange between the shards (the
// reference has no batch notion; instances are independent).  Host code only, written over the single-device C-ABI.
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/cfnmpc.h"

struct cfnmpc_multi
{
    int B = 0, N = 0, n = 0;
    std::vector<cfnmpc_batch *> h;
    std::vector<int> first, count, dev;
};

static thread_local std::string g_merr;
extern "C" const char *cfnmpc_multi_last_error(void) { return g_merr.c_str(); }
static int mfail(int code, const std::string &msg)
{
    g_merr = msg;
    return code;
}

// run fn(i) for every shard on its own host thread; first failure wins
template <class F>
static int for_each_shard(cfnmpc_multi *m, F fn)
{
    std::vector<int> rc(m->n, 0);
    std::vector<std::string> msg(m->n);
    std::vector<std::thread> th;
    for (int i = 0; i < m->n; i++)
        th.emplace_back([&, i]() {
            rc[i] = fn(i);
            if (rc[i] != CFNMPC_OK) msg[i] = cfnmpc_last_error();   // thread-local text of the worker
        });
    for (auto &t : th) t.join();
    for (int i = 0; i < m->n; i++)
        if (rc[i] != CFNMPC_OK) return mfail(rc[i], "device " + std::to_string(m->dev[i]) + ": " + msg[i]);
    return CFNMPC_OK;
}

// bytes per instance of a per-instance field (0: not a per-instance field)
static size_t per_instance_bytes(const cfnmpc_multi *m, const char *f, int stage_get)
{
    const size_t N = m->N;
    if (!strcmp(f, "x0") || !strcmp(f, "yref_e") || !strcmp(f, "W_e_batch")) return 13 * 8;
    if (!strcmp(f, "yref")) return N * 17 * 8;
    if (!strcmp(f, "x_all") || (!strcmp(f, "x") && !stage_get)) return (N + 1) * 13 * 8;
    if (!strcmp(f, "u_all") || (!strcmp(f, "u") && !stage_get)) return N * 4 * 8;
    if (!strcmp(f, "x")) return 13 * 8;
    if (!strcmp(f, "u")) return 4 * 8;
    if (!strcmp(f, "W_batch")) return 17 * 8;
    if (!strcmp(f, "lbu_batch") || !strcmp(f, "ubu_batch") || !strcmp(f, "lbu0_batch") || !strcmp(f, "ubu0_batch") || !strcmp(f, "twist") ||
        !strcmp(f, "res"))
        return 4 * 8;
    if (!strcmp(f, "setpoint") || !strcmp(f, "euler")) return 3 * 8;
    if (!strcmp(f, "status") || !strcmp(f, "qp_iter") || !strcmp(f, "qp_status") || !strcmp(f, "flags") || !strcmp(f, "policy") ||
        !strcmp(f, "traj_iter"))
        return 4;
    if (!strcmp(f, "motors")) return 4 * 4;
    return 0;
}

extern "C" int cfnmpc_multi_destroy(cfnmpc_multi *m)
{
    if (!m) return CFNMPC_OK;
    for (auto *b : m->h) if (b) cfnmpc_batch_destroy(b);
    delete m;
    return CFNMPC_OK;
}

extern "C" int cfnmpc_multi_create(int batch, int N, double Ts, int n_dev, const int *devices, cfnmpc_multi **out)
{
    if (!out) return mfail(CFNMPC_EINVAL, "cfnmpc_multi_create: out is NULL");
    *out = nullptr;
    if (n_dev < 1 || !devices || batch < n_dev) return mfail(CFNMPC_EINVAL, "cfnmpc_multi_create: need n_dev >= 1 devices and batch >= n_dev");
    cfnmpc_multi *m = new cfnmpc_multi();
    m->B = batch; m->N = N; m->n = n_dev;
    m->h.assign(n_dev, nullptr);
    const int q = batch / n_dev, r = batch % n_dev;
    for (int i = 0; i < n_dev; i++) {
        m->first.push_back(i * q + (i < r ? i : r));
        m->count.push_back(q + (i < r ? 1 : 0));
        m->dev.push_back(devices[i]);
    }
    const int rc = for_each_shard(m, [&](int i) { return cfnmpc_batch_create(m->count[i], N, Ts, m->dev[i], &m->h[i]); });
    if (rc != CFNMPC_OK) { cfnmpc_multi_destroy(m); return rc; }
    *out = m;
    return CFNMPC_OK;
}

extern "C" int cfnmpc_multi_shard(cfnmpc_multi *m, int i, cfnmpc_batch **h, int *device, int *first, int *count)
{
    if (!m || i < 0 || i >= m->n) return mfail(CFNMPC_EINVAL, "cfnmpc_multi_shard: no such shard");
    if (h) *h = m->h[i];
    if (device) *device = m->dev[i];
    if (first) *first = m->first[i];
    if (count) *count = m->count[i];
    return CFNMPC_OK;
}

extern "C" int cfnmpc_multi_num_shards(cfnmpc_multi *m) { return m ? m->n : 0; }

extern "C" int cfnmpc_multi_set(cfnmpc_multi *m, const char *field, const void *host_src)
{
    if (!m || !field || !host_src) return mfail(CFNMPC_EINVAL, "cfnmpc_multi_set: null argument");
    const size_t per = per_instance_bytes(m, field, 0);
    return for_each_shard(m, [&](int i) {
        const char *src = static_cast<const char *>(host_src) + (per ? per * (size_t) m->first[i] : 0);   // solver-wide fields: every shard
        const int rc = cfnmpc_batch_set(m->h[i], field, src, 0);
        return rc == CFNMPC_OK ? cfnmpc_batch_sync(m->h[i]) : rc;   // the caller may reuse its buffer on return
    });
}

extern "C" int cfnmpc_multi_set_option(cfnmpc_multi *m, const char *option, int value)
{
    if (!m || !option) return mfail(CFNMPC_EINVAL, "cfnmpc_multi_set_option: null argument");
    return for_each_shard(m, [&](int i) { return cfnmpc_batch_set_option(m->h[i], option, value); });
}

extern "C" int cfnmpc_multi_solve(cfnmpc_multi *m, int n_rti)
{
    if (!m) return mfail(CFNMPC_EINVAL, "null handle");
    return for_each_shard(m, [&](int i) { return cfnmpc_batch_solve(m->h[i], n_rti); });
}

extern "C" int cfnmpc_multi_solve_from_host(cfnmpc_multi *m, const double *x0, const double *yref, const double *yref_e, int n_chunks)
{
    if (!m || !x0 || !yref || !yref_e) return mfail(CFNMPC_EINVAL, "cfnmpc_multi_solve_from_host: null argument");
    const size_t N = m->N;
    return for_each_shard(m, [&](int i) {
        const size_t f = m->first[i];
        const int rc = cfnmpc_batch_solve_from_host(m->h[i], x0 + f * 13, yref + f * N * 17, yref_e + f * 13, n_chunks);
        return rc == CFNMPC_OK ? cfnmpc_batch_sync(m->h[i]) : rc;
    });
}

extern "C" int cfnmpc_multi_tick(cfnmpc_multi *m, int motors_from_u1)
{
    if (!m) return mfail(CFNMPC_EINVAL, "null handle");
    return for_each_shard(m, [&](int i) { return cfnmpc_batch_tick(m->h[i], motors_from_u1); });
}

extern "C" int cfnmpc_multi_set_trajectory(cfnmpc_multi *m, const double *table, int n_rows)
{
    if (!m || !table) return mfail(CFNMPC_EINVAL, "cfnmpc_multi_set_trajectory: null argument");
    return for_each_shard(m, [&](int i) {
        const int rc = cfnmpc_batch_set_trajectory(m->h[i], table, n_rows, 0);
        return rc == CFNMPC_OK ? cfnmpc_batch_sync(m->h[i]) : rc;
    });
}

extern "C" int cfnmpc_multi_sync(cfnmpc_multi *m)
{
    if (!m) return mfail(CFNMPC_EINVAL, "null handle");
    return for_each_shard(m, [&](int i) { return cfnmpc_batch_sync(m->h[i]); });
}

extern "C" int cfnmpc_multi_get(cfnmpc_multi *m, const char *field, int stage, void *host_dst)
{
    if (!m || !field || !host_dst) return mfail(CFNMPC_EINVAL, "cfnmpc_multi_get: null argument");
    const size_t per = per_instance_bytes(m, field, 1);
    if (!per) return mfail(CFNMPC_EINVAL, std::string("cfnmpc_multi_get: '") + field + "' is not a per-instance field");
    return for_each_shard(m, [&](int i) {
        return cfnmpc_batch_get(m->h[i], field, stage, static_cast<char *>(host_dst) + per * (size_t) m->first[i], 0);
    });
}

// device time of the last solve: the slowest shard (ms)
extern "C" int cfnmpc_multi_last_solve_ms(cfnmpc_multi *m, double *ms)
{
    if (!m || !ms) return mfail(CFNMPC_EINVAL, "null argument");
    std::vector<double> v(m->n, 0.0);
    const int rc = for_each_shard(m, [&](int i) { return cfnmpc_batch_last_solve_ms(m->h[i], &v[i]); });
    *ms = 0.0;
    for (double x : v) *ms = x > *ms ? x : *ms;
    return rc;
}
