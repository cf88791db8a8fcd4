/* Minimal stand-in for libf2c's f2c.h, written for this repo: only the
 * typedefs/macros the reference's f2c-generated AMOS sources need. */
#ifndef OB_STUB_F2C_H
#define OB_STUB_F2C_H
typedef long int integer;
typedef double doublereal;
typedef float real;
typedef long int logical;
#define TRUE_ (1)
#define FALSE_ (0)
#ifndef abs
#define abs(x) ((x) >= 0 ? (x) : -(x))
#endif
#ifndef min
#define min(a, b) ((a) <= (b) ? (a) : (b))
#endif
#ifndef max
#define max(a, b) ((a) >= (b) ? (a) : (b))
#endif
#endif
