/* Stand-in for the header cmake generates from srcAna/Types.in.h in a real build of the reference (configure_file with
 * every #cmakedefine off: serial, no Belos / MPI / ScaLAPACK).  Written for this repo so that some of the reference's
 * translation units compile where they lie (oracle/Makefile, target ref_path); it only has to provide the scalar type
 * names and the two container aliases those files use, here on top of the container-only <Eigen/Core> of this
 * directory. */
#ifndef OPTIMET_TYPES_H
#define OPTIMET_TYPES_H
#include <Eigen/Core>
#include <complex>
#include <cstddef>
#include <functional>
namespace optimet {
using t_real = double;                  /* reals */
using t_complex = std::complex<t_real>; /* complex numbers */
using t_int = int;                      /* signed indices */
using t_uint = std::size_t;             /* unsigned indices */
/* column-major dynamic containers */
template <class SCALAR = t_complex> using Matrix = Eigen::Matrix<SCALAR, Eigen::Dynamic, Eigen::Dynamic, Eigen::ColMajor>;
template <class SCALAR = t_complex> using Vector = Eigen::Matrix<SCALAR, Eigen::Dynamic, 1, Eigen::ColMajor>;
} // namespace optimet
#endif
