"""Golden vector for the ray trace: the Zemax wavefront map and annular Zernike listing that the reference's
own test holds (/root/reference/tests/test_opd.py:16-95; files tests/data/LSST_{WF,AZ}_v3.3_c3_f6_w3_M2_dx_100um.txt:
LSST Ver. 3.3 baseline design, M2 decentred by 100 um in x, field (1.121, 1.231) deg, 694 nm, 256 x 256 pupil grid).
Stored with the orientation the reference test compares in (flipped in y, Zemax's one-sample zero border dropped)."""
import os

import numpy as np

DATA = "/root/reference/tests/data/"
HERE = os.path.dirname(os.path.abspath(__file__))

with open(DATA + "LSST_WF_v3.3_c3_f6_w3_M2_dx_100um.txt", encoding="utf-16-le") as f:
    wf = np.genfromtxt(f, skip_header=16)
wf = np.flipud(wf)[1:, 1:]  # test_opd.py:80-81
with open(DATA + "LSST_AZ_v3.3_c3_f6_w3_M2_dx_100um.txt", encoding="utf-16-le") as f:
    zk = np.genfromtxt(f, skip_header=32, usecols=(2))
np.savez_compressed(os.path.join(HERE, "zemax_opd.npz"), opd_waves=wf, annular_zernike_waves=zk[:28],
                    wavelength_nm=694.0, thx_deg=1.121, thy_deg=1.231, m2_shift=np.array([100e-6, 0.0, 0.0]),
                    eps=0.612, projection="zemax")
print(wf.shape, np.count_nonzero(wf), zk[:4])
