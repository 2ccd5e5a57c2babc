// One translation unit per compiled N_rhs: nvcc -DBCG_N=<n> inst.cu
#ifndef BCG_N
#error "compile with -DBCG_N=<n>"
#endif
#include "ops.cuh"

#define BCG_CAT_(a, b) a##b
#define BCG_CAT(a, b) BCG_CAT_(a, b)

namespace bcg {
const OpsTable* BCG_CAT(get_ops_, BCG_N)() { return make_ops<BCG_N>(); }
}  // namespace bcg
