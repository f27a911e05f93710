import sys; sys.path.insert(0, '/root/repo')
import numpy as np
from argus_gui_b200 import ba
from oracle import ba_port
g = np.load('/root/repo/tests/golden/blockdiag_golden.npz')
p0 = g["M1_p0"]; M = 4
np.set_printoptions(linewidth=200, precision=6)
for it in (3, 4, 5, 6, 7, 8, 10):
    cams, info, ret = ba.solve_mot(g["M1_vmask"], p0[:64], p0[64:], g["M1_rot0"], g["M1_x"], 5, 4, itmax=it)
    pp, ip, _ = ba_port.port_mot(p0, g["M1_x"], g["M1_vmask"], g["M1_rot0"], 5, 4, itmax=it)
    print(it, "gpu ", info[1], info[2:5], info[5:])
    print(it, "port", ip[1], ip[2:5], ip[5:])
