// ob_multi.cu -- one process, several GPUs: a group of contexts driven by one host worker thread per GPU.
//
// The reference's solver::factory returns the serial solver unconditionally in non-MPI builds
// (srcAna/Solver.cpp:30-34), so a serial Optimet3D sees ONE solver object; this group lets that object use every GPU of
// the box.  It is written on top of the public C ABI only (ob_create, ob_comm_init, ob_run, ...): rank r of the group
// is exactly the rank r of a one-process-per-GPU launch, with the NCCL communicators created from threads of this
// process (ncclCommInitRank rendezvous between the workers).
#include "../../include/optimet_b200.h"
#include <functional>
#include <string>
#include <thread>
#include <vector>

struct ob_multi {
  std::vector<ob_ctx *> ctx;
  std::string err;
};

namespace {
// fn(rank) on every context concurrently (one host thread per GPU); first failing rank's message is kept
int for_all(ob_multi *m, std::function<int(int)> const &fn) {
  const int n = (int)m->ctx.size();
  std::vector<int> rc(n, 0);
  if(n == 1) {
    rc[0] = fn(0);
  } else {
    std::vector<std::thread> th;
    for(int r = 0; r < n; ++r)
      th.emplace_back([&, r] { rc[r] = fn(r); });
    for(auto &t : th)
      t.join();
  }
  for(int r = 0; r < n; ++r)
    if(rc[r]) {
      m->err = "rank " + std::to_string(r) + ": " + ob_last_error(m->ctx[r]);
      return rc[r];
    }
  return 0;
}
} // namespace

extern "C" {

int ob_create_multi(int ngpu, const int *devices, ob_multi **out) {
  if(!out)
    return 1;
  *out = nullptr;
  if(ngpu < 1 || !devices)
    return 1;
  ob_multi *m = new ob_multi();
  for(int r = 0; r < ngpu; ++r) {
    ob_ctx *c = nullptr;
    if(ob_create(devices[r], &c)) {
      for(ob_ctx *p : m->ctx)
        ob_destroy(p);
      delete m;
      return 1; // ob_last_error(NULL) carries the reason
    }
    m->ctx.push_back(c);
  }
  if(ngpu > 1) {
    char uid[128];
    if(ob_comm_unique_id(uid) || for_all(m, [&](int r) { return ob_comm_init(m->ctx[r], uid, r, ngpu); })) {
      for(ob_ctx *p : m->ctx)
        ob_destroy(p);
      delete m;
      return 1;
    }
  }
  *out = m;
  return 0;
}

void ob_destroy_multi(ob_multi *m) {
  if(!m)
    return;
  for(ob_ctx *p : m->ctx)
    ob_destroy(p);
  delete m;
}

int ob_multi_size(const ob_multi *m) { return m ? (int)m->ctx.size() : 0; }
ob_ctx *ob_multi_ctx(ob_multi *m, int rank) { return (m && rank >= 0 && rank < (int)m->ctx.size()) ? m->ctx[rank] : nullptr; }
const char *ob_multi_last_error(ob_multi *m) { return m ? m->err.c_str() : "null group"; }

int ob_multi_set_cluster(ob_multi *m, int nobj, const double *xyz_m, const double *radius_m, int nMax, int nMaxS) {
  if(!m)
    return 1;
  return for_all(m, [&](int r) { return ob_set_cluster(m->ctx[r], nobj, xyz_m, radius_m, nMax, nMaxS); });
}

int ob_multi_set_frequency(ob_multi *m, double omega, const double waveK[2], const double eps_b[2], const double mu_b[2],
                           const double *eps, const double *mu, const double *eps_SH, const double *mu_SH,
                           const double *ksippp, const double *ksiparppar, const double *gamma) {
  if(!m)
    return 1;
  return for_all(m, [&](int r) {
    return ob_set_frequency(m->ctx[r], omega, waveK, eps_b, mu_b, eps, mu, eps_SH, mu_SH, ksippp, ksiparppar, gamma);
  });
}

int ob_multi_set_incident(ob_multi *m, const double *a, const double *b) {
  if(!m)
    return 1;
  return for_all(m, [&](int r) { return ob_set_incident(m->ctx[r], a, b); });
}

int ob_multi_set_option(ob_multi *m, const char *name, double value) {
  if(!m)
    return 1;
  int rc = 0; // options include process-wide tuning knobs: set them one context after the other
  for(size_t r = 0; r < m->ctx.size() && !rc; ++r)
    rc = ob_set_option(m->ctx[r], name, value);
  if(rc)
    m->err = ob_last_error(m->ctx[0]);
  return rc;
}

int ob_multi_run(ob_multi *m, const ob_gmres_opts *opts, int do_sh, double *X_sca, double *X_int, double *X_sca_SH,
                 double *X_int_SH, double cs[5], int stats[2]) {
  if(!m)
    return 1;
  const int n = (int)m->ctx.size();
  std::vector<double> part((size_t)5 * n, 0.0);
  std::vector<int> st((size_t)2 * n, 0);
  const int rc = for_all(m, [&](int r) {
    return ob_run(m->ctx[r], opts, do_sh, r == 0 ? X_sca : nullptr, r == 0 ? X_int : nullptr, r == 0 ? X_sca_SH : nullptr,
                  r == 0 ? X_int_SH : nullptr, &part[(size_t)5 * r], &st[(size_t)2 * r]);
  });
  if(rc)
    return rc;
  if(cs)
    for(int k = 0; k < 5; ++k) { // per-rank partial sums over the rank's own particles
      cs[k] = 0.0;
      for(int r = 0; r < n; ++r)
        cs[k] += part[(size_t)5 * r + k];
    }
  if(stats) {
    stats[0] = st[0];
    stats[1] = st[1];
  }
  return 0;
}

} // extern "C"
