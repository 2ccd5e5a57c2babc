// N_rhs values compiled into the library (one inst.cu object each).
#pragma once
#ifndef BCG_FOR_EACH_N
#define BCG_FOR_EACH_N(X) X(1) X(2) X(3) X(4) X(6) X(8) X(12) X(16) X(32)
#endif
