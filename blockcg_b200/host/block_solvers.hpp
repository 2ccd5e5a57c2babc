// Host-side mirror of the reference's inc/block_solvers.hpp: identical signatures,
// defaults and return value (number of operator applications); X is overwritten,
// B, D and sigma are left untouched.  The whole iteration runs on the GPU
// (bcg_solve_* of include/blockcg_b200.h); the host only moves B in and X out.
#ifndef BLOCKCG_B200_HOST_BLOCK_SOLVERS_H
#define BLOCKCG_B200_HOST_BLOCK_SOLVERS_H
#include <algorithm>
#include <cassert>

#include "dirac_op.hpp"
#include "fields.hpp"

// inc/block_solvers.hpp:10-45
template <int N_rhs>
int BCG(block_fermion_field<N_rhs>& X, const block_fermion_field<N_rhs>& B, const dirac_op& D, double eps = 1.e-15,
        int max_iterations = 1e6) {
  bcg_ctx* c = D.bind<N_rhs>();
  bcg_solve_info info;
  bcg_host::check(c, bcg_solve_bcg(c, X.raw(), B.raw(), eps, max_iterations, &info), "bcg_solve_bcg");
  return info.iterations;
}

// inc/block_solvers.hpp:50-86
template <int N_rhs>
int BCGrQ(block_fermion_field<N_rhs>& X, const block_fermion_field<N_rhs>& B, const dirac_op& D, double eps = 1.e-15,
          int max_iterations = 1e6) {
  bcg_ctx* c = D.bind<N_rhs>();
  bcg_solve_info info;
  bcg_host::check(c, bcg_solve_bcgrq(c, X.raw(), B.raw(), eps, max_iterations, &info), "bcg_solve_bcgrq");
  return info.iterations;
}

// inc/block_solvers.hpp:91-185
template <int N_rhs>
int SBCGrQ(std::vector<block_fermion_field<N_rhs>>& X, const block_fermion_field<N_rhs>& B, const dirac_op& D,
           std::vector<double>& sigma, double eps = 1.e-15, double eps_shifts = 1.e-15, int max_iterations = 1e6) {
  assert(sigma.size() == X.size() && "number of shifts does not match number of solution vectors");
  assert(sigma[0] >= 0.0 && "shifts must be zero or positive");
  assert(std::is_sorted(sigma.begin(), sigma.end()) && "shifts must be in ascending order");
  const int n_shifts = static_cast<int>(sigma.size());
  bcg_ctx* c = D.bind<N_rhs>(n_shifts);
  std::vector<double*> xp(n_shifts);
  for (int s = 0; s < n_shifts; ++s) xp[s] = X[s].raw();
  bcg_solve_info info;
  bcg_host::check(c, bcg_solve_sbcgrq(c, xp.data(), B.raw(), sigma.data(), n_shifts, eps, eps_shifts, max_iterations,
                                      &info),
                  "bcg_solve_sbcgrq");
  return info.iterations;
}

#endif  // BLOCKCG_B200_HOST_BLOCK_SOLVERS_H
