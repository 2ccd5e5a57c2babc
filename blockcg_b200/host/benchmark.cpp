// The reference's benchmark program (benchmark.cpp) on the GPU path: same positional
// CLI (V mass eps [eps_shifts]), same N_rhs = 12 and shift list, same report lines
// (worst true relative residual per shift, iteration counts), plus wall-clock times,
// which the reference does not print.  SCG per column goes through the N=1 block path.
#include <chrono>
#include <cmath>
#include <iostream>

#include "block_solvers.hpp"
#include "standard_solvers.hpp"

constexpr int N_rhs = 12;

static double seconds() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char* argv[]) {
  std::vector<double> shifts = {0, 0, 1e-10, 1e-8, 1e-6, 1e-5, 1e-4, 1e-2, 1e-1};
  const int N_shifts = static_cast<int>(shifts.size());
  if (argc - 1 < 3) {
    std::cout << "usage: benchmark V mass eps [eps_shifts = 1e-15] [--no-scg]\n"
              << "e.g. ./benchmark 1024 0.01 1e-12" << std::endl;
    return 1;
  }
  const int V = static_cast<int>(atof(argv[1]));
  const double mass = atof(argv[2]);
  const double eps = atof(argv[3]);
  double eps_shifts = 1.e-15;
  bool run_scg = true;
  for (int i = 4; i < argc; ++i) {
    if (std::string(argv[i]) == "--no-scg")
      run_scg = false;
    else
      eps_shifts = atof(argv[i]);
  }

  dirac_op D(V, mass);
  block_fermion_field<N_rhs> B(V);
  B.setRandom();

  std::cout << "# Benchmark of SBCGrQ vs SCG solver: V = " << V << ", N_rhs = " << N_rhs << ", mass = " << mass
            << ", eps = " << eps << ", eps_shifts = " << eps_shifts << std::endl
            << std::endl;
  std::cout << "# Shifts:\t\t";
  for (double s : shifts) std::cout << std::scientific << s << "\t";
  std::cout << std::endl << std::endl;

  int iterSCG = 0;
  double tSCG = 0;
  if (run_scg) {
    std::vector<double> resSCG(N_shifts, 0.0);
    fermion_field b(V), Ax(V);
    std::vector<fermion_field> x(N_shifts, b);
    for (int i_rhs = 0; i_rhs < N_rhs; ++i_rhs) {
      for (int i_x = 0; i_x < V; ++i_x)
        for (int c = 0; c < N_f; ++c) b[i_x](c, 0) = B[i_x](c, i_rhs);
      double t0 = seconds();
      iterSCG += SCG(x, b, D, shifts, eps, eps_shifts);
      tSCG += seconds() - t0;
      const double b2 = b.real_dot(b);
      for (int s = 0; s < N_shifts; ++s) {
        D.op(Ax, x[s]);
        Ax.add(x[s], shifts[s]);
        Ax -= b;
        resSCG[s] = std::max(resSCG[s], std::sqrt(Ax.real_dot(Ax) / b2));
      }
    }
    std::cout << "# SCG residuals:\t";
    for (double r : resSCG) std::cout << std::scientific << r << "\t";
    std::cout << std::endl;
  }

  block_fermion_field<N_rhs> AX(V);
  std::vector<block_fermion_field<N_rhs>> X(N_shifts, B);
  double t0 = seconds();
  const int iterSBCGrQ = N_rhs * SBCGrQ(X, B, D, shifts, eps, eps_shifts);
  const double tS = seconds() - t0;
  std::cout << "# SBCGrQ residuals:\t";
  block_matrix<N_rhs> b2 = B.hermitian_dot(B);
  for (int s = 0; s < N_shifts; ++s) {
    D.op(AX, X[s]);
    AX.add(X[s], shifts[s]);
    AX -= B;
    block_matrix<N_rhs> r2 = AX.hermitian_dot(AX);
    double res2 = 0;
    for (int i = 0; i < N_rhs; ++i) res2 = std::max(res2, r2(i, i).real() / b2(i, i).real());
    std::cout << std::scientific << std::sqrt(res2) << "\t";
  }
  std::cout << std::endl << std::endl;
  if (run_scg) std::cout << "# SCG_iterations:\t" << iterSCG << std::endl;
  std::cout << "# SBCGrQ_iterations:\t" << iterSBCGrQ << std::endl;
  if (run_scg) std::cout << "# SCG_seconds:\t\t" << std::fixed << tSCG << std::endl;
  std::cout << "# SBCGrQ_seconds:\t" << std::fixed << tS << std::endl;
  return 0;
}
