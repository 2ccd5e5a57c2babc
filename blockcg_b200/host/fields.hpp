// Host-side mirror of the reference's inc/fields.hpp on top of the blockcg_b200 C-ABI.
//
// Same public names, argument meaning and memory layout as the reference
// (block_fermion<N> = 3 x N complex<double> column-major, 48N bytes; a field is a
// contiguous array of them => [V][N][3] complex128; block_matrix<N> = N x N
// column-major), but no Eigen: the element types are small PODs that offer the
// handful of accessors the reference's own drivers use (operator()(r,c), col(i),
// diagonal entries, data(), setZero, setRandom, Identity, adjoint).
//
// Every field-sized operation of the hot path (add, rescale_add, hermitian_dot,
// multiply_upper_triangular_inverse_RHS, thinQR) runs on the GPU through
// include/blockcg_b200.h; nothing here falls back to CPU arithmetic.  Errors of
// the C-ABI surface as std::runtime_error (the reference has no error channel).
#ifndef BLOCKCG_B200_HOST_FIELDS_H
#define BLOCKCG_B200_HOST_FIELDS_H
#include <complex>
#include <cstdlib>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/blockcg_b200.h"

constexpr int N_f = 3;  // inc/fields.hpp:18

namespace bcg_host {

inline void check(bcg_ctx* c, int rc, const char* what) {
  if (rc != BCG_OK)
    throw std::runtime_error(std::string(what) + ": " + (c ? bcg_last_error(c) : "no context") + " (status " +
                             std::to_string(rc) + ")");
}

// Lattice geometry of the operator that fields are currently bound to: the reference's operator
// is a 1-D chain of V sites (dims empty); the 4-D extension carries (L0, L1, L2, L3).
inline std::vector<long long>& current_dims() {
  static std::vector<long long> d;
  return d;
}

// one device context per (V, N, geometry); grown when more shifts are requested
inline bcg_ctx* context(int V, int N, int n_shifts = 1) {
  struct Entry {
    bcg_ctx* ctx;
    int S;
  };
  static std::map<std::pair<std::pair<int, int>, std::vector<long long>>, Entry> cache;
  const std::vector<long long>& dims = current_dims();
  auto key = std::make_pair(std::make_pair(V, N), dims);
  auto it = cache.find(key);
  if (it != cache.end() && it->second.S >= n_shifts) return it->second.ctx;
  if (it != cache.end()) {
    bcg_ctx_destroy(it->second.ctx);
    cache.erase(it);
  }
  bcg_ctx* c = nullptr;
  int rc;
  if (dims.size() == 4) {
    const int64_t d[4] = {dims[0], dims[1], dims[2], dims[3]};
    rc = bcg_ctx_create_4d(&c, d, N, n_shifts, 0, 0, 1);
  } else {
    rc = bcg_ctx_create(&c, V, N, n_shifts, 0, 0, 1);
  }
  if (rc != BCG_OK) {
    std::string msg = c ? bcg_last_error(c) : "bcg_ctx_create failed";
    bcg_ctx_destroy(c);
    throw std::runtime_error(msg);
  }
  cache[key] = Entry{c, n_shifts};
  return c;
}

// Eigen's Matrix::setRandom as the reference uses it (Core/MathFunctions.h:618-628,
// 715-727): uniform in [-1,1], imaginary part drawn first (gcc argument order).
inline std::complex<double> random_complex() {
  auto rnd = []() { return -1.0 + 2.0 * double(std::rand()) / double(RAND_MAX); };
  double im = rnd();
  double re = rnd();
  return {re, im};
}

template <int R, int C>
struct small_matrix {
  std::complex<double> a[R * C];  // column-major
  std::complex<double>& operator()(int r, int c) { return a[r + R * c]; }
  const std::complex<double>& operator()(int r, int c) const { return a[r + R * c]; }
  std::complex<double>* data() { return a; }
  const std::complex<double>* data() const { return a; }
  std::complex<double>* col(int c) { return a + R * c; }
  const std::complex<double>* col(int c) const { return a + R * c; }
  static constexpr int rows() { return R; }
  static constexpr int cols() { return C; }
  void setZero() {
    for (auto& v : a) v = 0.0;
  }
  void setRandom() {
    for (auto& v : a) v = random_complex();
  }
  static small_matrix Identity() {
    small_matrix m;
    m.setZero();
    for (int i = 0; i < (R < C ? R : C); ++i) m(i, i) = 1.0;
    return m;
  }
  small_matrix<C, R> adjoint() const {
    small_matrix<C, R> m;
    for (int r = 0; r < R; ++r)
      for (int c = 0; c < C; ++c) m(c, r) = std::conj((*this)(r, c));
    return m;
  }
  // the expression forms the reference's callers write: -alpha, (beta * rho.adjoint()).eval(), 2.0 * M
  small_matrix operator-() const {
    small_matrix m;
    for (int i = 0; i < R * C; ++i) m.a[i] = -a[i];
    return m;
  }
  template <int C2>
  small_matrix<R, C2> operator*(const small_matrix<C, C2>& o) const {
    small_matrix<R, C2> m;
    m.setZero();
    for (int c = 0; c < C2; ++c)
      for (int k = 0; k < C; ++k)
        for (int r = 0; r < R; ++r) m(r, c) += (*this)(r, k) * o(k, c);
    return m;
  }
  small_matrix operator*(double s) const {
    small_matrix m;
    for (int i = 0; i < R * C; ++i) m.a[i] = a[i] * s;
    return m;
  }
  friend small_matrix operator*(double s, const small_matrix& m) { return m * s; }
  small_matrix operator+(const small_matrix& o) const { return small_matrix(*this) += o; }
  small_matrix operator-(const small_matrix& o) const { return small_matrix(*this) -= o; }
  const small_matrix& eval() const { return *this; }
  small_matrix& operator+=(const small_matrix& o) {
    for (int i = 0; i < R * C; ++i) a[i] += o.a[i];
    return *this;
  }
  small_matrix& operator-=(const small_matrix& o) {
    for (int i = 0; i < R * C; ++i) a[i] -= o.a[i];
    return *this;
  }
};

// RAII device field
struct dev_field {
  bcg_ctx* c;
  int h;
  dev_field(bcg_ctx* ctx, const void* host = nullptr) : c(ctx), h(-1) {
    check(c, bcg_field_alloc(c, &h), "bcg_field_alloc");
    if (host) check(c, bcg_field_upload(c, h, static_cast<const double*>(host)), "bcg_field_upload");
  }
  ~dev_field() {
    if (h >= 0) bcg_field_free(c, h);
  }
  void download(void* host) { check(c, bcg_field_download(c, h, static_cast<double*>(host)), "bcg_field_download"); }
  dev_field(const dev_field&) = delete;
  dev_field& operator=(const dev_field&) = delete;
};

}  // namespace bcg_host

template <int N_rhs>
using block_fermion = bcg_host::small_matrix<N_f, N_rhs>;  // inc/fields.hpp:19-20
typedef block_fermion<1> fermion;
template <int N_rhs>
using block_matrix = bcg_host::small_matrix<N_rhs, N_rhs>;  // inc/fields.hpp:22-23

template <int N_rhs>
class block_fermion_field {  // inc/fields.hpp:25-147
 protected:
  std::vector<block_fermion<N_rhs>> data_;

 public:
  int V;

  explicit block_fermion_field(int V) : V(V) { data_.resize(V); }
  block_fermion<N_rhs>& operator[](int i) { return data_[i]; }
  const block_fermion<N_rhs>& operator[](int i) const { return data_[i]; }
  double* raw() { return reinterpret_cast<double*>(data_.data()); }
  const double* raw() const { return reinterpret_cast<const double*>(data_.data()); }

  block_fermion_field<N_rhs>& operator+=(const block_fermion_field<N_rhs>& rhs) {
    for (int ix = 0; ix < V; ++ix) data_[ix] += rhs[ix];
    return *this;
  }
  block_fermion_field<N_rhs>& operator-=(const block_fermion_field<N_rhs>& rhs) {
    for (int ix = 0; ix < V; ++ix) data_[ix] -= rhs[ix];
    return *this;
  }
  void setZero() {
    for (int ix = 0; ix < V; ++ix) data_[ix].setZero();
  }
  void setRandom() {
    for (int ix = 0; ix < V; ++ix) data_[ix].setRandom();
  }

  // this <- this + rhs * M   (inc/fields.hpp:70-77)
  block_fermion_field<N_rhs>& add(const block_fermion_field<N_rhs>& rhs, const block_matrix<N_rhs>& M) {
    bcg_ctx* c = bcg_host::context(V, N_rhs);
    bcg_host::dev_field d(c, raw()), s(c, rhs.raw());
    bcg_host::check(c, bcg_add(c, d.h, s.h, reinterpret_cast<const double*>(M.data())), "bcg_add");
    d.download(raw());
    return *this;
  }
  // this <- this + rhs * s   (real scalar overload, block_solvers.hpp:136)
  block_fermion_field<N_rhs>& add(const block_fermion_field<N_rhs>& rhs, double s) {
    bcg_ctx* c = bcg_host::context(V, N_rhs);
    bcg_host::dev_field d(c, raw()), r(c, rhs.raw());
    bcg_host::check(c, bcg_add_scalar(c, d.h, r.h, s), "bcg_add_scalar");
    d.download(raw());
    return *this;
  }
  // this <- this * L + rhs * r   (inc/fields.hpp:79-90)
  block_fermion_field<N_rhs>& rescale_add(const block_matrix<N_rhs>& L, const block_fermion_field<N_rhs>& rhs,
                                          double r) {
    bcg_ctx* c = bcg_host::context(V, N_rhs);
    bcg_host::dev_field d(c, raw()), s(c, rhs.raw());
    bcg_host::check(c, bcg_rescale_add(c, d.h, reinterpret_cast<const double*>(L.data()), s.h, r), "bcg_rescale_add");
    d.download(raw());
    return *this;
  }
  // scalar left multiplier, as dirac_op.hpp:42 (lhs.rescale_add(-1.0, rhs, mass * mass)) and CG / SCG use it
  block_fermion_field<N_rhs>& rescale_add(double l, const block_fermion_field<N_rhs>& rhs, double r) {
    return rescale_add(block_matrix<N_rhs>::Identity() * l, rhs, r);
  }
  // Re(this . rhs) for N_rhs == 1 (inc/fields.hpp:93-99): diagonal of the 1x1 Gram
  double real_dot(const block_fermion_field<1>& rhs) const {
    static_assert(N_rhs == 1, "real_dot is defined for fermion_field");
    return hermitian_dot(rhs)(0, 0).real();
  }
  // R_ij = this_i . rhs_j, lower triangle accumulated, upper mirrored (inc/fields.hpp:103-122)
  block_matrix<N_rhs> hermitian_dot(const block_fermion_field<N_rhs>& rhs) const {
    bcg_ctx* c = bcg_host::context(V, N_rhs);
    block_matrix<N_rhs> R;
    bcg_host::dev_field a(c, raw());
    if (&rhs == this) {
      bcg_host::check(c, bcg_gram(c, a.h, a.h, reinterpret_cast<double*>(R.data())), "bcg_gram");
    } else {
      bcg_host::dev_field b(c, rhs.raw());
      bcg_host::check(c, bcg_gram(c, a.h, b.h, reinterpret_cast<double*>(R.data())), "bcg_gram");
    }
    return R;
  }
  // this <- this R^-1, R upper triangular (inc/fields.hpp:125-136)
  block_fermion_field<N_rhs>& multiply_upper_triangular_inverse_RHS(const block_matrix<N_rhs>& R) {
    bcg_ctx* c = bcg_host::context(V, N_rhs);
    bcg_host::dev_field q(c, raw());
    bcg_host::check(c, bcg_trsm(c, q.h, reinterpret_cast<const double*>(R.data())), "bcg_trsm");
    q.download(raw());
    return *this;
  }
  // thin QR by Cholesky (inc/fields.hpp:140-146)
  block_fermion_field<N_rhs>& thinQR(block_matrix<N_rhs>& R) {
    bcg_ctx* c = bcg_host::context(V, N_rhs);
    bcg_host::dev_field q(c, raw());
    bcg_host::check(c, bcg_thinqr(c, q.h, reinterpret_cast<double*>(R.data())), "bcg_thinqr");
    q.download(raw());
    return *this;
  }
};
typedef block_fermion_field<1> fermion_field;

#endif  // BLOCKCG_B200_HOST_FIELDS_H
