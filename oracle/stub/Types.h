/* Stand-in for the reference's cmake-generated Types.h (srcAna/Types.in.h with every #cmakedefine off: serial build,
 * no Belos / MPI / ScaLAPACK), written for this repo so that a few of the reference's own translation units can be
 * compiled where they lie (oracle/Makefile, target ref_coupling).  Same typedefs and aliases as Types.in.h:33-52. */
#ifndef OPTIMET_TYPES_H
#define OPTIMET_TYPES_H
#include <complex>
#include <functional>
#include <Eigen/Core>
namespace optimet {
typedef int t_int;
typedef std::size_t t_uint;
typedef double t_real;
typedef std::complex<t_real> t_complex;
template <class T = t_complex> using Vector = Eigen::Matrix<T, Eigen::Dynamic, 1, Eigen::ColMajor>;
template <class T = t_complex> using Matrix = Eigen::Matrix<T, Eigen::Dynamic, Eigen::Dynamic, Eigen::ColMajor>;
}
#endif
