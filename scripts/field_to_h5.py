#!/usr/bin/env python
"""Convert a <case>_FF.field / <case>_SH.field file written by optimet3d_b200 into the reference's HDF5 layout
(Output.cpp:25-54, OutputGrid.cpp:41-92): groups Field_E and Field_H, each with X, Y, Z/{real, imag} and ABS/abs as
[nx][ny][nz] doubles.  Needs h5py (not part of the build image, which is why the C++ side writes the raw form)."""
import sys

import numpy as np


def read_field(path):
    raw = open(path, "rb").read()
    head, body = raw.split(b"\n", 1)
    tok = head.decode().split()
    assert tok[0] == "OPTIMET_B200_FIELD" and tok[1] == "1"
    nx, ny, nz = (int(v) for v in tok[2:5])
    names = tok[5:]
    data = np.frombuffer(body, dtype="<f8").reshape(len(names), nx, ny, nz)
    return dict(zip(names, data))


def main():
    import h5py
    src, dst = sys.argv[1], sys.argv[2]
    with h5py.File(dst, "w") as f:
        for name, arr in read_field(src).items():
            f.create_dataset(name, data=arr)


if __name__ == "__main__":
    main()
