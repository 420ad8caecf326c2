"""Picks the strongest available checker for a cfg string (TEST INFRASTRUCTURE):
  1. "reference": the unmodified reference driven through oracle/_ref (RefDrv), or
  2. "port": the oracle's own C restatement (oracle/oracle_*.c) on the compiled material.
Used only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg."""
import os

from _libs import RefDrv, have_refdrv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _RefOracle:
    kind = "reference"

    def __init__(self, cfg):
        self.r = RefDrv(cfg)

    def xs_iso(self, ekin):
        return self.r.xs_iso(ekin)

    def sample_iso(self, ekin, seed, first_index=0):
        return self.r.sample_iso(ekin, seed=seed, first_index=first_index)

    def xs(self, ekin, ux, uy, uz):
        return self.r.xs(ekin, ux, uy, uz)

    def sample(self, ekin, ux, uy, uz, seed, first_index=0):
        return self.r.sample(ekin, ux, uy, uz, seed=seed, first_index=first_index)


def material_path(cfg):
    import ctypes as C
    from ncrystal_b200 import _lib
    buf = C.create_string_buffer(512)
    _lib.lib().ncb200_cfg_to_filestem(cfg.encode(), buf, 512)
    return os.path.join(ROOT, "ncrystal_b200", "data", buf.value.decode() + ".ncb")


def oracle_for(cfg, prefer=None):
    if prefer != "port" and have_refdrv():
        return _RefOracle(cfg)
    from _oracle_port import PortOracle
    return PortOracle(open(material_path(cfg), "rb").read())
