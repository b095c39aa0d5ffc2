/* Boost.Math is absent from this image.  The reference needs it only for the
 * `dipolar_rep` Jastrow (dmc/wavefunction/jastrows/jastrowDipolar.h:31-62), which no
 * BASELINE config uses.  This stand-in lets the unmodified reference main.cpp /
 * jastrowDipolar.cpp compile into oracle/_ref; calling it aborts ("parity unpinned" for
 * that one Jastrow).  TEST INFRASTRUCTURE, not product code. */
#ifndef QMC4_ORACLE_BOOST_BESSEL_PRIME_STUB
#define QMC4_ORACLE_BOOST_BESSEL_PRIME_STUB
#include <cstdio>
#include <cstdlib>
#endif
