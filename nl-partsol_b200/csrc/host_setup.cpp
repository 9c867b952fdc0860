// host_setup.cpp -- scalable host-side setup for the B200 engine.
//
// O(N) construction of the nodal adjacency the stepped path reads, reproducing the
// reference's linked-list CHAIN ORDERS exactly so that neighbour lists stay bit-identical:
//   element connectivity chains = file order reversed      (Nodes/Read-GID-Mesh.c:400-408)
//   NodeNeighbour[i]            = elements in DESCENDING id (InOutFun/Read_GramsBox.c:293-330)
//   NodalLocality_0[i] (1 ring) = union, push-front         (Read_GramsBox.c:367-398, Matlib/ChainOp.c:275-293)
//   NodalLocality[i]  (2 rings) = ring search, push-front   (Read_GramsBox.c:401-456)
//   h_avg[i]                    = mean 1-ring distance      (Read_GramsBox.c:460-507)
//   DeltaX                      = min element edge          (Read_GramsBox.c:510-565, Q4.c:457, H8.c:643)
// The reference's own construction is O(Nn*Ne) (get_sourrounding_elements) and cannot reach
// the 10^6-10^7 particle configurations.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/nlps_b200.h"

namespace {
struct Locality {
  std::vector<int> r1p, r1i, r2p, r2i;
  std::vector<double> h_avg;
  double dx;
};

void build(int d, int nn, int ne, int nne, const int* conn, const double* X, Locality& L) {
  // node -> elements, descending element id
  std::vector<int> nnp(nn + 1, 0);
  for (long e = 0; e < (long)ne * nne; e++) nnp[conn[e] + 1]++;
  for (int i = 0; i < nn; i++) nnp[i + 1] += nnp[i];
  std::vector<int> nni(nnp[nn]), fill(nnp.begin(), nnp.end() - 1);
  for (int e = ne - 1; e >= 0; e--)
    for (int k = 0; k < nne; k++) nni[fill[conn[(size_t)e * nne + k]]++] = e;
  // ring 1 (two passes: sizes, fill), thread-parallel with private stamps
  L.r1p.assign(nn + 1, 0);
  std::vector<int> tmp1((size_t)nn * 0);
  std::vector<std::vector<int>> rows;  // not used for large meshes; we do count+fill instead
  auto ring1 = [&](int I, std::vector<int>& stamp, int tag, std::vector<int>& out) {
    out.clear();
    for (int q = nnp[I]; q < nnp[I + 1]; q++) {
      int e = nni[q];
      for (int k = nne - 1; k >= 0; k--) {
        int v = conn[(size_t)e * nne + k];
        if (stamp[v] != tag) { stamp[v] = tag; out.push_back(v); }
      }
    }
  };
#pragma omp parallel
  {
    std::vector<int> stamp(nn, -1), out;
#pragma omp for schedule(static)
    for (int i = 0; i < nn; i++) {
      ring1(i, stamp, i, out);
      L.r1p[i + 1] = (int)out.size();
    }
  }
  for (int i = 0; i < nn; i++) L.r1p[i + 1] += L.r1p[i];
  L.r1i.resize(L.r1p[nn]);
  L.h_avg.assign(nn, 0.0);
#pragma omp parallel
  {
    std::vector<int> stamp(nn, -1), out;
#pragma omp for schedule(static)
    for (int i = 0; i < nn; i++) {
      ring1(i, stamp, i, out);
      int n1 = (int)out.size(), o = L.r1p[i];
      for (int k = n1 - 1; k >= 0; k--) L.r1i[o++] = out[k];  // chain order = reverse discovery
      double avg = 0.0;
      int cnt = 0;
      for (int q = L.r1p[i]; q < L.r1p[i + 1]; q++) {
        int B = L.r1i[q];
        if (B == i) continue;
        double aux = 0.0;
        for (int k = 0; k < d; k++) {
          double h = X[(size_t)B * d + k] - X[(size_t)i * d + k];
          aux += h * h;
        }
        avg += pow(aux, 0.5);
        cnt++;
      }
      L.h_avg[i] = avg / (double)cnt;
    }
  }
  // ring 2 from ring-1 rows (already in chain order)
  L.r2p.assign(nn + 1, 0);
  auto ring2 = [&](int I, std::vector<int>& stamp, int tag, std::vector<int>& S, std::vector<int>& search,
                   std::vector<int>& fresh) {
    S.clear();
    search.assign(1, I);
    for (int ring = 0; ring < 2; ring++) {
      fresh.clear();
      for (int s : search)
        for (int q = L.r1p[s]; q < L.r1p[s + 1]; q++) {
          int v = L.r1i[q];
          if (stamp[v] != tag) { stamp[v] = tag; S.push_back(v); fresh.push_back(v); }
        }
      search.assign(fresh.rbegin(), fresh.rend());
    }
  };
#pragma omp parallel
  {
    std::vector<int> stamp(nn, -1), S, search, fresh;
#pragma omp for schedule(static)
    for (int i = 0; i < nn; i++) {
      ring2(i, stamp, i, S, search, fresh);
      L.r2p[i + 1] = (int)S.size();
    }
  }
  for (int i = 0; i < nn; i++) L.r2p[i + 1] += L.r2p[i];
  L.r2i.resize(L.r2p[nn]);
#pragma omp parallel
  {
    std::vector<int> stamp(nn, -1), S, search, fresh;
#pragma omp for schedule(static)
    for (int i = 0; i < nn; i++) {
      ring2(i, stamp, i, S, search, fresh);
      int o = L.r2p[i];
      for (int k = (int)S.size() - 1; k >= 0; k--) L.r2i[o++] = S[k];
    }
  }
  // DeltaX
  double mn = 10e16;
  static const int ed8[12][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {4, 5}, {5, 6}, {6, 7}, {7, 4}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
  for (int e = 0; e < ne; e++) {
    int nedge = (nne == 4) ? 4 : 12;
    for (int k = 0; k < nedge; k++) {
      int a, b;
      if (nne == 4) { a = conn[(size_t)e * 4 + (3 - k)]; b = conn[(size_t)e * 4 + (3 - ((k + 1) % 4))]; }
      else { a = conn[(size_t)e * 8 + ed8[k][0]]; b = conn[(size_t)e * 8 + ed8[k][1]]; }
      double s = 0.0;
      for (int l = 0; l < d; l++) { double h = X[(size_t)b * d + l] - X[(size_t)a * d + l]; s += h * h; }
      double len = sqrt(s);
      if (len < mn) mn = len;
    }
  }
  L.dx = mn;
}
}  // namespace

extern "C" int nlps_b200_build_locality(int ndim, int n_nodes, int n_elems, int nodes_per_elem,
                                        const int* connectivity, const double* coords, int* ring1_ptr,
                                        int* ring1_idx, int* ring2_ptr, int* ring2_idx, double* h_avg,
                                        double* delta_x) {
  if ((ndim != 2 && ndim != 3) || (nodes_per_elem != 4 && nodes_per_elem != 8)) return 1;
  for (long i = 0; i < (long)n_elems * nodes_per_elem; i++)
    if (connectivity[i] < 0 || connectivity[i] >= n_nodes) return 1;
  // two-call protocol: the second call reuses the result of the first.  The cache belongs to the calling THREAD
  // (several host threads may set up different meshes at once) and is keyed on the sizes, the connectivity pointer
  // and a digest of the connectivity, not on the pointer alone (a freed and reallocated buffer can have the same address)
  static thread_local Locality cache;
  static thread_local const int* cache_key = nullptr;
  static thread_local unsigned long long cache_digest = 0;
  unsigned long long digest = 1469598103934665603ull ^ (unsigned long long)n_nodes ^ ((unsigned long long)n_elems << 32);
  for (long i = 0; i < (long)n_elems * nodes_per_elem; i += 97) digest = (digest ^ (unsigned)connectivity[i]) * 1099511628211ull;
  if (cache_key != connectivity || cache_digest != digest || (int)cache.r1p.size() != n_nodes + 1) {
    build(ndim, n_nodes, n_elems, nodes_per_elem, connectivity, coords, cache);
    cache_key = connectivity;
    cache_digest = digest;
  }
  memcpy(ring1_ptr, cache.r1p.data(), sizeof(int) * (n_nodes + 1));
  memcpy(ring2_ptr, cache.r2p.data(), sizeof(int) * (n_nodes + 1));
  if (h_avg) memcpy(h_avg, cache.h_avg.data(), sizeof(double) * n_nodes);
  if (delta_x) *delta_x = cache.dx;
  if (ring1_idx && ring2_idx) {
    memcpy(ring1_idx, cache.r1i.data(), sizeof(int) * cache.r1i.size());
    memcpy(ring2_idx, cache.r2i.data(), sizeof(int) * cache.r2i.size());
    cache = Locality();
    cache_key = nullptr;
  }
  return 0;
}

// ---------------------------------------------------------------------------
// Spatial-slab planning.  The reference has no multi-process path; these are the host
// decisions every slab must take identically: the slab axis, the cuts, who owns a particle (by the
// coordinate of its closest node I0) and which nodes two neighbour slabs exchange.
extern "C" int nlps_b200_slab_cuts(const nlps_mesh* mesh, int n, const int* I0, int world, int axis, int* axis_out,
                                   double* cuts_out) {
  const int d = mesh->ndim;
  if (world < 1 || n < 1 || axis >= d) return 1;
  for (int p = 0; p < n; p++)
    if (I0[p] < 0 || I0[p] >= mesh->n_nodes) return 1;
  if (axis < 0) {  // longest extent of the cloud of closest nodes
    double best = -1.0;
    for (int k = 0; k < d; k++) {
      double lo = 1e300, hi = -1e300;
      for (int p = 0; p < n; p++) {
        const double c = mesh->coords[(size_t)I0[p] * d + k];
        lo = std::min(lo, c);
        hi = std::max(hi, c);
      }
      if (hi - lo > best) { best = hi - lo; axis = k; }
    }
  }
  if (axis_out) *axis_out = axis;
  std::vector<double> c(n);
  for (int p = 0; p < n; p++) c[p] = mesh->coords[(size_t)I0[p] * d + axis];
  std::sort(c.begin(), c.end());
  double prev_cut = -1e300;
  for (int g = 1; g < world; g++) {
    size_t k = (size_t)((double)g * n / world);
    if (k >= (size_t)n) k = n - 1;
    double q = c[k];
    // first layer of this slab = q; the cut sits midway to the previous distinct layer
    auto it = std::lower_bound(c.begin(), c.end(), q);
    double below = (it == c.begin()) ? q - mesh->delta_x : *(it - 1);
    double cut = 0.5 * (q + below);
    if (!(cut > prev_cut)) return 2;  // more slabs than node layers with particles
    cuts_out[g - 1] = cut;
    prev_cut = cut;
  }
  return 0;
}

extern "C" int nlps_b200_slab_owner(const nlps_mesh* mesh, int axis, int world, const double* cuts, int I0) {
  const double c = mesh->coords[(size_t)I0 * mesh->ndim + axis];
  int g = 0;
  while (g < world - 1 && c >= cuts[g]) g++;
  return g;
}

extern "C" int nlps_b200_slab_halo_nodes(const nlps_mesh* mesh, int axis, double cut, int band_cells, int* ids) {
  const double w = band_cells * mesh->delta_x * (1.0 + 1e-9);
  int n = 0;
  for (int i = 0; i < mesh->n_nodes; i++) {
    const double c = mesh->coords[(size_t)i * mesh->ndim + axis];
    if (fabs(c - cut) <= w) {
      if (ids) ids[n] = i;
      n++;
    }
  }
  return n;
}
