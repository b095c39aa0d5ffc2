/* Boost.Math is absent from this image.  The reference needs it only for the `dipolar_rep` Jastrow
 * (dmc/wavefunction/jastrows/jastrowDipolar.h:31-62), which no BASELINE config uses.  This stand-in lets the unmodified
 * reference main.cpp / jastrowDipolar.cpp compile into oracle/_ref and run: cyl_bessel_k for integer orders comes from
 * oracle/bessel_k.h (trapezoidal rule on the integral representation + forward recurrence; pinned on scipy.special.kv,
 * not on Boost — "parity unpinned" against the reference's own binary for that one Jastrow).
 * TEST INFRASTRUCTURE, not product code. */
#ifndef QMC4_ORACLE_BOOST_BESSEL_STUB
#define QMC4_ORACLE_BOOST_BESSEL_STUB
#include "../../../../bessel_k.h"
namespace boost { namespace math {
template <class T1, class T2> inline double cyl_bessel_k(T1 n, T2 x) { return orc_cyl_bessel_k((int)n, (double)x); }
}}
#endif
