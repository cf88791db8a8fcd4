/* Stand-in for <hdf5.h>: the handle types OutputGrid.h names in its declarations (HDF5 is absent from the image).
 * Only used to COMPILE the reference's Solver.h for tests/test_adaptor_compiles.py; nothing here is ever called. */
#pragma once
typedef long long hid_t;
typedef int herr_t;
typedef unsigned long long hsize_t;
