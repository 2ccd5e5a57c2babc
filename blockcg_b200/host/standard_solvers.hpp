// Host-side mirror of the reference's inc/standard_solvers.hpp (CG, SCG for one
// right-hand side) -- a "next" row of the scope table, provided through the block
// path: for N_rhs = 1 block CG is CG and the multishift block solver solves the same
// shifted systems as SCG (src/standard_solvers.cpp:3-95).  Signatures, stopping
// rule (|r|/|b| < eps on the lowest shift) and return value are the reference's;
// the iterates are those of BCG<1> / SBCGrQ<1>, i.e. equal up to rounding, not
// bit-for-bit, because the scalar recurrences are evaluated as 1x1 block recurrences.
#ifndef BLOCKCG_B200_HOST_STANDARD_SOLVERS_H
#define BLOCKCG_B200_HOST_STANDARD_SOLVERS_H
#include "block_solvers.hpp"

inline int CG(fermion_field& x, const fermion_field& b, const dirac_op& D, double eps = 1.e-15,
              int max_iterations = 1e6) {
  return BCG<1>(x, b, D, eps, max_iterations);
}

inline int SCG(std::vector<fermion_field>& x, const fermion_field& b, const dirac_op& D, std::vector<double>& sigma,
               double eps = 1.e-15, double eps_shifts = 1.e-15, int max_iterations = 1e6) {
  return SBCGrQ<1>(x, b, D, sigma, eps, eps_shifts, max_iterations);
}

#endif  // BLOCKCG_B200_HOST_STANDARD_SOLVERS_H
