/* double-double ("quad") helpers with the macro names of SCIP's scip/dbldblarith.h; own implementation
 * (Knuth two-sum / FMA two-product). Only used by tightenRowCoefs in the reference's sdpi.c. */
#ifndef SHIM_SCIP_DBLDBLARITH_H
#define SHIM_SCIP_DBLDBLARITH_H
#include <math.h>
#define QUAD_HI(x)  x ## hi
#define QUAD_LO(x)  x ## lo
#define QUAD(x)     QUAD_HI(x), QUAD_LO(x)
#define QUAD_TO_DBL(x) ( QUAD_HI(x) + QUAD_LO(x) )
#define QUAD_ASSIGN(a, c) do { QUAD_HI(a) = (c); QUAD_LO(a) = 0.0; } while( 0 )
#define QUAD_ASSIGN_Q(a, b) do { QUAD_HI(a) = QUAD_HI(b); QUAD_LO(a) = QUAD_LO(b); } while( 0 )
static inline void shimTwoSum(double a, double b, double* s, double* e)
{
   volatile double sum = a + b; volatile double bb = sum - a; *e = (a - (sum - bb)) + (b - bb); *s = sum;
}
static inline void shimRenorm(double hi, double lo, double* rh, double* rl)
{
   volatile double s = hi + lo; *rl = lo - (s - hi); *rh = s;
}
#define SCIPquadprecSumDD(r, a, b) do { double _s, _e; shimTwoSum((a), (b), &_s, &_e); QUAD_HI(r) = _s; QUAD_LO(r) = _e; } while( 0 )
#define SCIPquadprecSumQD(r, a, b) do { double _s, _e; shimTwoSum(QUAD_HI(a), (b), &_s, &_e); _e += QUAD_LO(a); \
      shimRenorm(_s, _e, &QUAD_HI(r), &QUAD_LO(r)); } while( 0 )
#define SCIPquadprecSumQQ(r, a, b) do { double _s, _e, _t, _f; shimTwoSum(QUAD_HI(a), QUAD_HI(b), &_s, &_e); \
      shimTwoSum(QUAD_LO(a), QUAD_LO(b), &_t, &_f); _e += _t; shimRenorm(_s, _e, &_s, &_e); _e += _f; \
      shimRenorm(_s, _e, &QUAD_HI(r), &QUAD_LO(r)); } while( 0 )
#define SCIPquadprecProdQD(r, a, b) do { double _p = QUAD_HI(a) * (b); double _e = fma(QUAD_HI(a), (b), -_p); \
      _e += QUAD_LO(a) * (b); shimRenorm(_p, _e, &QUAD_HI(r), &QUAD_LO(r)); } while( 0 )
#endif
