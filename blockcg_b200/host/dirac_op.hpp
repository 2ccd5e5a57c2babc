// Host-side mirror of the reference's inc/dirac_op.hpp: same public surface (V, mass,
// ctor, op<N>) plus the links accessor the reference lacks (its U is private,
// inc/dirac_op.hpp:9-11).  op() runs the sm_100a block stencil through the C-ABI.
#ifndef BLOCKCG_B200_HOST_DIRAC_OP_H
#define BLOCKCG_B200_HOST_DIRAC_OP_H
#include <array>

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
  // 4-D extension (NOT in the reference): L0 x L1 x L2 x L3 periodic lattice, four links per site
  // stored [site][mu]; D v[x] = 1/2 sum_mu (U_mu[x] v[x+mu] - U_mu[x-mu]^dag v[x-mu]), op = m^2 - D^2.
  dirac_op(const std::array<int, 4>& L, double mass) : U(4 * static_cast<size_t>(L[0]) * L[1] * L[2] * L[3]),
                                                       V(L[0] * L[1] * L[2] * L[3]), mass(mass), dims(L.begin(), L.end()) {
    for (auto& u : U) u.setRandom();
  }
  std::vector<long long> dims;  // empty: the reference's 1-D chain
  const std::complex<double>* links() const { return U[0].data(); }
  std::complex<double>* links() { return U[0].data(); }

  // make this operator the one the (V, N) device context applies
  template <int N_rhs>
  bcg_ctx* bind(int n_shifts = 1) const {
    bcg_host::current_dims() = dims;  // fields created from now on live on this operator's lattice
    bcg_ctx* c = bcg_host::context(V, N_rhs, n_shifts);
    if (dims.size() == 4)
      bcg_host::check(c, bcg_set_links_4d(c, reinterpret_cast<const double*>(links()), mass), "bcg_set_links_4d");
    else
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
