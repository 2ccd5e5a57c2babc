// Host-side mirror of the reference's inc/dirac_op.hpp: same public surface (V, mass,
// ctor, op<N>) plus the links accessor the reference lacks (its U is private,
// inc/dirac_op.hpp:9-11).  op() runs the sm_100a block stencil through the C-ABI.
#ifndef BLOCKCG_B200_HOST_DIRAC_OP_H
#define BLOCKCG_B200_HOST_DIRAC_OP_H
#include "fields.hpp"

class dirac_op {
 private:
  using gauge = bcg_host::small_matrix<N_f, N_f>;
  std::vector<gauge> U;

 public:
  int V;
  double mass;

  // random 3x3 complex "gauge links", entries uniform in [-1,1]+i[-1,1], drawn from
  // std::rand() exactly as the reference does (inc/dirac_op.hpp:24-32)
  explicit dirac_op(int V, double mass = 0.1) : U(V), V(V), mass(mass) {
    for (int ix = 0; ix < V; ++ix) U[ix].setRandom();
  }
  const std::complex<double>* links() const { return U[0].data(); }
  std::complex<double>* links() { return U[0].data(); }

  // make this operator the one the (V, N) device context applies
  template <int N_rhs>
  bcg_ctx* bind(int n_shifts = 1) const {
    bcg_ctx* c = bcg_host::context(V, N_rhs, n_shifts);
    bcg_host::check(c, bcg_set_links(c, reinterpret_cast<const double*>(links()), mass), "bcg_set_links");
    return c;
  }

  // lhs = (m^2 - D^2) rhs   (inc/dirac_op.hpp:36-43)
  template <int N_rhs>
  void op(block_fermion_field<N_rhs>& lhs, const block_fermion_field<N_rhs>& rhs) const {
    bcg_ctx* c = bind<N_rhs>();
    bcg_host::dev_field in(c, rhs.raw()), out(c);
    bcg_host::check(c, bcg_op(c, out.h, in.h, 0.0, nullptr), "bcg_op");
    out.download(lhs.raw());
  }
};

#endif  // BLOCKCG_B200_HOST_DIRAC_OP_H
