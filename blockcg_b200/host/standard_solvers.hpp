// Host-side mirror of the reference's inc/standard_solvers.hpp: CG and SCG for one right-hand
// side, same signatures, defaults, stopping rules and return value (number of operator
// applications) as src/standard_solvers.cpp:3-32 and :34-95.  The scalar recurrences
// (alpha, beta and, per shift, zeta / theta) run on the GPU as the reference writes them
// (bcg_solve_cg / bcg_solve_scg of include/blockcg_b200.h: scalars in the device-side control
// block, no N x N algebra); the host only moves b in and the solutions out.
#ifndef BLOCKCG_B200_HOST_STANDARD_SOLVERS_H
#define BLOCKCG_B200_HOST_STANDARD_SOLVERS_H
#include <algorithm>
#include <cassert>

#include "dirac_op.hpp"
#include "fields.hpp"

// src/standard_solvers.cpp:3-32
inline int CG(fermion_field& x, const fermion_field& b, const dirac_op& D, double eps = 1.e-15,
              int max_iterations = 1e6) {
  bcg_ctx* c = D.bind<1>();
  bcg_solve_info info;
  bcg_host::check(c, bcg_solve_cg(c, x.raw(), b.raw(), eps, max_iterations, &info), "bcg_solve_cg");
  return info.iterations;
}

// src/standard_solvers.cpp:34-95
inline int SCG(std::vector<fermion_field>& x, const fermion_field& b, const dirac_op& D, std::vector<double>& sigma,
               double eps = 1.e-15, double eps_shifts = 1.e-15, int max_iterations = 1e6) {
  assert(sigma.size() == x.size() && "number of shifts does not match number of solution vectors");
  assert(sigma[0] >= 0.0 && "shifts must be zero or positive");
  assert(std::is_sorted(sigma.begin(), sigma.end()) && "shifts must be in ascending order");
  const int n_shifts = static_cast<int>(sigma.size());
  bcg_ctx* c = D.bind<1>(n_shifts);
  std::vector<double*> xp(n_shifts);
  for (int s = 0; s < n_shifts; ++s) xp[s] = x[s].raw();
  bcg_solve_info info;
  bcg_host::check(c, bcg_solve_scg(c, xp.data(), b.raw(), sigma.data(), n_shifts, eps, eps_shifts, max_iterations,
                                   &info),
                  "bcg_solve_scg");
  return info.iterations;
}

#endif  // BLOCKCG_B200_HOST_STANDARD_SOLVERS_H
