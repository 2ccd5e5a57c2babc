// Acceptance tests of the reference (test/solvers.cpp: V=128, mass 0.5, eps 1e-10,
// N_rhs=3, shifts {0,0.01,0.1,0.2,0.9}; "true residual < 2*eps" per rhs per shift)
// run against the GPU path through the drop-in host headers.  No test framework:
// every check prints one line, the exit code is the number of failures.
#include <cmath>
#include <cstdio>

#include "block_solvers.hpp"
#include "standard_solvers.hpp"

static int V = 128;
static double mass = 0.5;
static double stopping_criterion = 1.e-10;
constexpr int N_rhs = 3;
static std::vector<double> shifts = {0.0, 0.01, 0.10, 0.20, 0.9};

static int failures = 0, checks = 0;
static void require_lt(double value, double bound, const char* what, int a = -1, int b = -1) {
  ++checks;
  const bool ok = value < bound;  // false for NaN, as REQUIRE(residual < ...) would be
  if (!ok) ++failures;
  std::printf("%s %-8s", ok ? "ok  " : "FAIL", what);
  if (a >= 0) std::printf(" shift %d", a);
  if (b >= 0) std::printf(" rhs %d", b);
  std::printf("  residual %.3e < %.1e\n", value, bound);
}

template <int N>
static void check_block(const char* name, const block_fermion_field<N>& X, const block_fermion_field<N>& B,
                        const dirac_op& D, double shift, int i_shift) {
  block_fermion_field<N> AX(V);
  D.op(AX, X);
  if (shift != 0.0) AX.add(X, shift);
  AX -= B;
  block_matrix<N> r2 = AX.hermitian_dot(AX);
  block_matrix<N> b2 = B.hermitian_dot(B);
  for (int i = 0; i < N; ++i)
    require_lt(std::sqrt(r2(i, i).real() / b2(i, i).real()), 2 * stopping_criterion, name, i_shift, i);
}

int main() {
  const int n_shifts = static_cast<int>(shifts.size());
  {  // CG
    fermion_field x(V), b(V);
    dirac_op D(V, mass);
    b.setRandom();
    int it = CG(x, b, D, stopping_criterion);
    std::printf("CG: %d iterations\n", it);
    check_block<1>("CG", x, b, D, 0.0, -1);
  }
  {  // SCG
    fermion_field b(V);
    dirac_op D(V, mass);
    std::vector<fermion_field> x(n_shifts, b);
    b.setRandom();
    int it = SCG(x, b, D, shifts, stopping_criterion);
    std::printf("SCG: %d iterations\n", it);
    for (int s = 0; s < n_shifts; ++s) check_block<1>("SCG", x[s], b, D, shifts[s], s);
  }
  {  // BCG
    block_fermion_field<N_rhs> X(V), B(V);
    dirac_op D(V, mass);
    B.setRandom();
    int it = BCG(X, B, D, stopping_criterion);
    std::printf("BCG: %d iterations\n", it);
    check_block<N_rhs>("BCG", X, B, D, 0.0, -1);
  }
  {  // BCGrQ
    block_fermion_field<N_rhs> X(V), B(V);
    dirac_op D(V, mass);
    B.setRandom();
    int it = BCGrQ(X, B, D, stopping_criterion);
    std::printf("BCGrQ: %d iterations\n", it);
    check_block<N_rhs>("BCGrQ", X, B, D, 0.0, -1);
  }
  {  // SBCGrQ
    block_fermion_field<N_rhs> B(V);
    dirac_op D(V, mass);
    std::vector<block_fermion_field<N_rhs>> X(n_shifts, B);
    B.setRandom();
    int it = SBCGrQ(X, B, D, shifts, stopping_criterion);
    std::printf("SBCGrQ: %d iterations\n", it);
    for (int s = 0; s < n_shifts; ++s) check_block<N_rhs>("SBCGrQ", X[s], B, D, shifts[s], s);
  }
  std::printf("%s (%d assertions, %d failed)\n", failures ? "FAILED" : "All tests passed", checks, failures);
  const int ref_failures = failures;
  {  // 4-D extension of the operator (not part of the reference's suite): SBCGrQ on a 6 x 4 x 4 x 8 lattice
    checks = failures = 0;
    const std::array<int, 4> L = {6, 4, 4, 8};
    V = L[0] * L[1] * L[2] * L[3];
    dirac_op D(L, mass);
    block_fermion_field<N_rhs> B(V);
    std::vector<block_fermion_field<N_rhs>> X(n_shifts, B);
    B.setRandom();
    int it = SBCGrQ(X, B, D, shifts, stopping_criterion);
    std::printf("4-D SBCGrQ: %d iterations\n", it);
    for (int s = 0; s < n_shifts; ++s) check_block<N_rhs>("SBCGrQ4", X[s], B, D, shifts[s], s);
    std::printf("4-D extension: %s (%d assertions, %d failed)\n", failures ? "FAILED" : "passed", checks, failures);
  }
  return ref_failures + failures;
}
