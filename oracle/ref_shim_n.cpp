// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Per-N translation unit of the reference shim: compiles the UNMODIFIED
// reference headers where they lie (/root/reference/inc, passed with -I) for
// one compile-time N_rhs (-DNRHS=<n>) and exports plain-C entry points that
// take raw (re,im)-interleaved double buffers in the reference's own layout:
//   field  : [V][N][3] complex128  (site-major; 3xN column-major per site,
//            reference inc/fields.hpp:18-30)
//   links  : [V][3][3] complex128  (column-major 3x3, inc/dirac_op.hpp:10-11)
//   matrix : N x N complex128 column-major (inc/fields.hpp:22-23)
// Nothing here restates the algorithm: every function forwards to the
// reference's own templates, so outputs of this library ARE the reference.
//
// dirac_op::U is private (inc/dirac_op.hpp:9-11); this TU (and only this TU)
// opens it with the usual preprocessor trick so tests can feed identical
// links to the reference and to the CUDA path.
#include <chrono>
#include <cstdlib>
#include <cstring>
// Pull every system / Eigen header in first so the access hack below touches
// only the reference's own two class definitions.
#include <algorithm>
#include <complex>
#include <vector>
#include "Eigen3/Eigen/Dense"
#include "Eigen3/Eigen/StdVector"
#define private public
#include "dirac_op.hpp"
#undef private
#include "block_solvers.hpp"
#if NRHS == 1
#include "standard_solvers.hpp"  // CG / SCG: compiled from the reference's own src/standard_solvers.cpp (Makefile)
#endif

#ifndef NRHS
#error "compile with -DNRHS=<n>"
#endif

#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, NRHS)

namespace {
constexpr int N = NRHS;
using field_t = block_fermion_field<N>;
using mat_t = block_matrix<N>;
using cplx = std::complex<double>;

void load_field(field_t& f, const double* p) {
  std::memcpy(static_cast<void*>(&f[0](0, 0)), p, sizeof(double) * 2 * 3 * N * f.V);
}
void store_field(const field_t& f, double* p) {
  std::memcpy(p, static_cast<const void*>(&f[0](0, 0)), sizeof(double) * 2 * 3 * N * f.V);
}
void load_mat(mat_t& m, const double* p) { std::memcpy(static_cast<void*>(m.data()), p, sizeof(double) * 2 * N * N); }
void store_mat(const mat_t& m, double* p) { std::memcpy(p, static_cast<const void*>(m.data()), sizeof(double) * 2 * N * N); }
void load_links(dirac_op& D, const double* U) {
  std::memcpy(static_cast<void*>(D.U[0].data()), U, sizeof(double) * 2 * 9 * D.V);
}
double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
}  // namespace

extern "C" {

// Inputs exactly as benchmark.cpp:36-40 makes them: links first, then B.
void FN(ref_make_inputs)(int V, unsigned seed, double* U_out, double* B_out) {
  std::srand(seed);
  dirac_op D(V, 0.1);
  field_t B(V);
  B.setRandom();
  std::memcpy(U_out, static_cast<const void*>(D.U[0].data()), sizeof(double) * 2 * 9 * V);
  store_field(B, B_out);
}

double FN(ref_op)(int V, double mass, const double* U, const double* in, double* out, int reps) {
  dirac_op D(V, mass);
  load_links(D, U);
  field_t x(V), y(V);
  load_field(x, in);
  double t0 = now();
  for (int r = 0; r < reps; ++r) D.op(y, x);
  double t1 = now();
  store_field(y, out);
  return (t1 - t0) / reps;
}

double FN(ref_hermitian_dot)(int V, const double* a, const double* b, double* out, int reps) {
  field_t x(V), y(V);
  load_field(x, a);
  load_field(y, b);
  mat_t R;
  double t0 = now();
  for (int r = 0; r < reps; ++r) R = x.hermitian_dot(y);
  double t1 = now();
  store_mat(R, out);
  return (t1 - t0) / reps;
}

void FN(ref_add)(int V, double* dst, const double* src, const double* M) {
  field_t x(V), y(V);
  load_field(x, dst);
  load_field(y, src);
  mat_t m;
  load_mat(m, M);
  x.add(y, m);
  store_field(x, dst);
}

void FN(ref_add_scalar)(int V, double* dst, const double* src, double s) {
  field_t x(V), y(V);
  load_field(x, dst);
  load_field(y, src);
  x.add(y, s);
  store_field(x, dst);
}

void FN(ref_rescale_add)(int V, double* dst, const double* L, const double* src, double r) {
  field_t x(V), y(V);
  load_field(x, dst);
  load_field(y, src);
  mat_t m;
  load_mat(m, L);
  x.rescale_add(m, y, r);
  store_field(x, dst);
}

void FN(ref_thinQR)(int V, double* q, double* R_out) {
  field_t x(V);
  load_field(x, q);
  mat_t R;
  x.thinQR(R);
  store_field(x, q);
  store_mat(R, R_out);
}

void FN(ref_fullpivlu_inverse)(const double* A, double* out) {
  mat_t a, r;
  load_mat(a, A);
  r = a.fullPivLu().solve(mat_t::Identity());
  store_mat(r, out);
}

void FN(ref_fullpivlu_solve)(const double* A, const double* B, double* out) {
  mat_t a, b, r;
  load_mat(a, A);
  load_mat(b, B);
  r = a.fullPivLu().solve(b);
  store_mat(r, out);
}

void FN(ref_llt_upper)(const double* A, double* out) {
  mat_t a, r;
  load_mat(a, A);
  r = a.llt().matrixL().adjoint();
  store_mat(r, out);
}

int FN(ref_BCG)(int V, double mass, const double* U, const double* B, double* X, double eps, int max_it,
                double* seconds) {
  dirac_op D(V, mass);
  load_links(D, U);
  field_t b(V), x(V);
  load_field(b, B);
  double t0 = now();
  int it = BCG<N>(x, b, D, eps, max_it);
  *seconds = now() - t0;
  store_field(x, X);
  return it;
}

int FN(ref_BCGrQ)(int V, double mass, const double* U, const double* B, double* X, double eps, int max_it,
                  double* seconds) {
  dirac_op D(V, mass);
  load_links(D, U);
  field_t b(V), x(V);
  load_field(b, B);
  double t0 = now();
  int it = BCGrQ<N>(x, b, D, eps, max_it);
  *seconds = now() - t0;
  store_field(x, X);
  return it;
}

// X: [S][V][N][3]
int FN(ref_SBCGrQ)(int V, double mass, const double* U, const double* B, double* X, const double* sigma,
                   int n_shifts, double eps, double eps_shifts, int max_it, double* seconds) {
  dirac_op D(V, mass);
  load_links(D, U);
  field_t b(V);
  load_field(b, B);
  std::vector<field_t> x(n_shifts, b);
  std::vector<double> sig(sigma, sigma + n_shifts);
  double t0 = now();
  int it = SBCGrQ<N>(x, b, D, sig, eps, eps_shifts, max_it);
  *seconds = now() - t0;
  for (int s = 0; s < n_shifts; ++s) store_field(x[s], X + static_cast<size_t>(s) * 2 * 3 * N * V);
  return it;
}

#if NRHS == 1
// CG / SCG exist for one right-hand side only (inc/standard_solvers.hpp:10-20)
int ref_CG_1(int V, double mass, const double* U, const double* B, double* X, double eps, int max_it,
             double* seconds) {
  dirac_op D(V, mass);
  load_links(D, U);
  field_t b(V), x(V);
  load_field(b, B);
  double t0 = now();
  int it = CG(x, b, D, eps, max_it);
  *seconds = now() - t0;
  store_field(x, X);
  return it;
}
int ref_SCG_1(int V, double mass, const double* U, const double* B, double* X, const double* sigma, int n_shifts,
              double eps, double eps_shifts, int max_it, double* seconds) {
  dirac_op D(V, mass);
  load_links(D, U);
  field_t b(V);
  load_field(b, B);
  std::vector<field_t> x(n_shifts, b);
  std::vector<double> sig(sigma, sigma + n_shifts);
  double t0 = now();
  int it = SCG(x, b, D, sig, eps, eps_shifts, max_it);
  *seconds = now() - t0;
  for (int s = 0; s < n_shifts; ++s) store_field(x[s], X + static_cast<size_t>(s) * 2 * 3 * V);
  return it;
}
#endif

}  // extern "C"
