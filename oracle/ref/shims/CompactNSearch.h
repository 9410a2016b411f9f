// TEST INFRASTRUCTURE ONLY — shim for InteractiveComputerGraphics/CompactNSearch
// (@3f11ece16a419fc1cc5795d6aa87cb7fe6b86960, pinned in CMake/NeighborhoodSearch.cmake:32-39; not vendored in
// /root/reference).  Restates the library's published behaviour for the calls the reference makes
// (Simulation.cpp:211-215, 746, 755, 848-898; Simulation.h:398-403, 520-543; FluidModel.cpp:228-230, 348, 363-374;
// BoundaryModel_Akinci2012.cpp:374, 396-413; Emitter.cpp:196):
//   * fixed-radius search between point sets, one neighbour list per (set, other set, point);
//   * a point is a neighbour iff the squared distance, accumulated as dx*dx, += dy*dy, += dz*dz, is < r*r (strict);
//     a point is never its own neighbour;
//   * an activation table says which (searching set, searched set) pairs are computed; add_point_set /
//     set_active(i, search, find) / set_active(i, j, b) / set_active(b) edit it like upstream's ActivationTable;
//   * z_sort() computes a Morton-order permutation per dynamic point set, sort_field() applies it.
// Neighbour lists come back in ascending index order (upstream: hash-grid traversal order; only the summation
// order of the callers depends on it).  The parity claim at this boundary is "unpinned": the reference holds no
// test or golden vector for neighbour sets, so this shim is the definition both sides are compared against.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <vector>

namespace CompactNSearch {
#ifdef USE_DOUBLE
using Real = double;
#else
using Real = float;
#endif

class NeighborhoodSearch;

class PointSet {
 public:
  std::size_t n_neighbors(unsigned int point_set, unsigned int i) const { return m_neighbors[point_set][i].size(); }
  unsigned int neighbor(unsigned int point_set, unsigned int i, unsigned int k) const { return m_neighbors[point_set][i][k]; }
  const std::vector<unsigned int> &neighbor_list(unsigned int point_set, unsigned int i) const { return m_neighbors[point_set][i]; }
  std::size_t n_points() const { return m_n; }
  bool is_dynamic() const { return m_dynamic; }
  void set_dynamic(bool v) { m_dynamic = v; }
  void *get_user_data() { return m_user_data; }
  Real const *point(unsigned int i) const { return &m_x[3 * i]; }
  template <typename T> void sort_field(T *lst) const {
    if (m_sort_table.empty()) return;
    std::vector<T> tmp(lst, lst + m_sort_table.size());
    for (std::size_t i = 0; i < m_sort_table.size(); i++) lst[i] = tmp[m_sort_table[i]];
  }

 private:
  friend class NeighborhoodSearch;
  PointSet(Real const *x, std::size_t n, bool dynamic, void *user_data) : m_x(x), m_n(n), m_dynamic(dynamic), m_user_data(user_data) {}
  void resize(Real const *x, std::size_t n) {
    m_x = x;
    m_n = n;
  }
  Real const *m_x;
  std::size_t m_n;
  bool m_dynamic;
  void *m_user_data;
  std::vector<std::vector<std::vector<unsigned int>>> m_neighbors;  // [other set][point][k]
  std::vector<unsigned int> m_sort_table;
};

class NeighborhoodSearch {
 public:
  NeighborhoodSearch(Real r, bool erase_empty_cells = false) : m_r(r), m_r2(r * r) { (void)erase_empty_cells; }
  PointSet const &point_set(unsigned int i) const { return m_sets[i]; }
  PointSet &point_set(unsigned int i) { return m_sets[i]; }
  std::size_t n_point_sets() const { return m_sets.size(); }
  std::vector<PointSet> const &point_sets() const { return m_sets; }
  std::vector<PointSet> &point_sets() { return m_sets; }
  Real radius() const { return m_r; }
  void set_radius(Real r) {
    m_r = r;
    m_r2 = r * r;
  }
  void resize_point_set(unsigned int i, Real const *x, std::size_t n) { m_sets[i].resize(x, n); }
  unsigned int add_point_set(Real const *x, std::size_t n, bool is_dynamic = true, bool search_neighbors = true, bool find_neighbors = true,
                             void *user_data = nullptr) {
    m_sets.push_back(PointSet(x, n, is_dynamic, user_data));
    const std::size_t size = m_table.size();
    for (std::size_t i = 0; i < size; i++) m_table[i].push_back(find_neighbors);
    m_table.push_back(std::vector<unsigned char>(size + 1, search_neighbors));
    m_table[size][size] = search_neighbors && find_neighbors;
    return (unsigned int)m_sets.size() - 1;
  }
  void set_active(unsigned int i, unsigned int j, bool active) { m_table[i][j] = active; }
  void set_active(unsigned int i, bool search_neighbors = true, bool find_neighbors = true) {
    const std::size_t size = m_table.size();
    for (std::size_t k = 0; k < size; k++) {
      m_table[k][i] = find_neighbors;
      m_table[i][k] = search_neighbors;
    }
    m_table[i][i] = search_neighbors && find_neighbors;
  }
  void set_active(bool active) {
    for (auto &row : m_table) std::fill(row.begin(), row.end(), (unsigned char)active);
  }
  bool is_active(unsigned int i, unsigned int j) const { return m_table[i][j] != 0; }
  void update_point_sets() {}

  void find_neighbors(bool points_changed = true) {
    (void)points_changed;
    const unsigned int ns = (unsigned int)m_sets.size();
    // sizes
    for (unsigned int a = 0; a < ns; a++) {
      m_sets[a].m_neighbors.resize(ns);
      for (unsigned int b = 0; b < ns; b++) {
        auto &lists = m_sets[a].m_neighbors[b];
        lists.resize(m_sets[a].m_n);
        for (auto &l : lists) l.clear();
      }
    }
    // grids per searched set (only those some active pair needs)
    std::vector<Grid> grids(ns);
    for (unsigned int b = 0; b < ns; b++) {
      bool needed = false;
      for (unsigned int a = 0; a < ns; a++) needed = needed || m_table[a][b];
      if (needed) build_grid(m_sets[b], grids[b]);
    }
    for (unsigned int a = 0; a < ns; a++)
      for (unsigned int b = 0; b < ns; b++) {
        if (!m_table[a][b]) continue;
        const PointSet &A = m_sets[a];
        const PointSet &B = m_sets[b];
        const Grid &g = grids[b];
        auto &lists = m_sets[a].m_neighbors[b];
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < (long long)A.m_n; i++) {
          const Real *xi = A.point((unsigned int)i);
          std::vector<unsigned int> &out = lists[i];
          const long long cx = cell_coord(xi[0]), cy = cell_coord(xi[1]), cz = cell_coord(xi[2]);
          for (long long dz = -1; dz <= 1; dz++)
            for (long long dy = -1; dy <= 1; dy++)
              for (long long dx = -1; dx <= 1; dx++) {
                const uint64_t key = cell_key(cx + dx, cy + dy, cz + dz);
                auto lo = std::lower_bound(g.keys.begin(), g.keys.end(), key);
                for (std::size_t p = (std::size_t)(lo - g.keys.begin()); p < g.keys.size() && g.keys[p] == key; p++) {
                  const unsigned int j = g.index[p];
                  if (a == b && j == (unsigned int)i) continue;
                  const Real *xj = B.point(j);
                  Real t = xi[0] - xj[0];
                  Real l2 = t * t;
                  t = xi[1] - xj[1];
                  l2 += t * t;
                  t = xi[2] - xj[2];
                  l2 += t * t;
                  if (l2 < m_r2) out.push_back(j);
                }
              }
          std::sort(out.begin(), out.end());
        }
      }
  }

  void z_sort() {
    for (auto &ps : m_sets) {
      ps.m_sort_table.clear();
      if (!ps.m_dynamic) continue;
      const std::size_t n = ps.m_n;
      std::vector<uint64_t> code(n);
      for (std::size_t i = 0; i < n; i++) {
        const Real *x = ps.point((unsigned int)i);
        code[i] = morton(cell_coord(x[0]), cell_coord(x[1]), cell_coord(x[2]));
      }
      ps.m_sort_table.resize(n);
      std::iota(ps.m_sort_table.begin(), ps.m_sort_table.end(), 0u);
      std::stable_sort(ps.m_sort_table.begin(), ps.m_sort_table.end(), [&](unsigned int a, unsigned int b) { return code[a] < code[b]; });
    }
  }

 private:
  struct Grid {
    std::vector<uint64_t> keys;        // sorted cell keys
    std::vector<unsigned int> index;   // point of each entry
  };
  long long cell_coord(Real v) const { return (long long)std::floor(v / m_r); }
  static uint64_t cell_key(long long x, long long y, long long z) {
    const uint64_t B = 1ull << 20;
    return ((uint64_t)(x + (long long)B) << 42) | ((uint64_t)(y + (long long)B) << 21) | (uint64_t)(z + (long long)B);
  }
  static uint64_t spread3(uint64_t v) {
    uint64_t o = 0;
    for (int b = 0; b < 21; b++) o |= ((v >> b) & 1ull) << (3 * b);
    return o;
  }
  static uint64_t morton(long long x, long long y, long long z) {
    const long long B = 1ll << 20;
    return spread3((uint64_t)(x + B)) | (spread3((uint64_t)(y + B)) << 1) | (spread3((uint64_t)(z + B)) << 2);
  }
  void build_grid(const PointSet &ps, Grid &g) const {
    const std::size_t n = ps.m_n;
    std::vector<std::pair<uint64_t, unsigned int>> e(n);
    for (std::size_t i = 0; i < n; i++) {
      const Real *x = ps.point((unsigned int)i);
      e[i] = {cell_key(cell_coord(x[0]), cell_coord(x[1]), cell_coord(x[2])), (unsigned int)i};
    }
    std::sort(e.begin(), e.end());
    g.keys.resize(n);
    g.index.resize(n);
    for (std::size_t i = 0; i < n; i++) {
      g.keys[i] = e[i].first;
      g.index[i] = e[i].second;
    }
  }
  Real m_r, m_r2;
  std::vector<PointSet> m_sets;
  std::vector<std::vector<unsigned char>> m_table;
};
}  // namespace CompactNSearch
