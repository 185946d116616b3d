#!/usr/bin/env python3
"""tools/reference_dump/dump_to_npz.py -- OSCDUMP1 file (written by osc_dump.h inside the
REFERENCE controller) -> .npz consumed by tests/test_reference_vectors.py.

usage: dump_to_npz.py in.oscdump out.npz --preset NAME [--adaptive-rho-interval K]
"""
import argparse
import struct

import numpy as np


def read_dump(path):
    raw = open(path, "rb").read()
    if raw[:8] != b"OSCDUMP1":
        raise ValueError(f"{path}: not an OSCDUMP1 file")
    nv, nu, nc, ns, n, m = struct.unpack_from("<6i", raw, 8)
    s = 6 * ns
    off = 8 + 24
    sizes = [("M", nv * nv), ("C", nv), ("J", s * nv), ("bias", s), ("targets", ns * 6), ("mask", nc)]
    rec_in = sum(k for _, k in sizes) * 8
    init, steps = None, []
    while off + 4 + rec_in <= len(raw):
        (kind,) = struct.unpack_from("<i", raw, off)
        off += 4
        d = {}
        for name, k in sizes:
            d[name] = np.frombuffer(raw, "<f8", k, off).copy()
            off += 8 * k
        if kind == 0:
            init = d
            continue
        if off + 8 * (n + m) + 8 > len(raw):
            break  # truncated last record
        d["solution"] = np.frombuffer(raw, "<f8", n, off).copy()
        off += 8 * n
        d["dual"] = np.frombuffer(raw, "<f8", m, off).copy()
        off += 8 * m
        d["exit_code"], d["iterations"] = struct.unpack_from("<2i", raw, off)
        off += 8
        steps.append(d)
    if init is None:
        raise ValueError("no Init record (kind 0): set_up_optimization was not recorded")
    return dict(nv=nv, nu=nu, nc=nc, ns=ns, n=n, m=m), init, steps


def to_npz(shape, init, steps, preset, interval):
    nv, ns, nc = shape["nv"], shape["ns"], shape["nc"]
    s = 6 * ns
    shp = dict(M=(nv, nv), C=(nv,), J=(s, nv), bias=(s,), targets=(ns, 6), mask=(nc,))
    out = {"preset": np.array(preset), "adaptive_rho_interval": np.array(int(interval)),
           "shape": np.array([shape[k] for k in ("nv", "nu", "nc", "ns", "n", "m")], np.int32)}
    for k, sh in shp.items():
        out["init_" + k] = init[k].reshape(sh)
        out[k] = np.stack([d[k].reshape(sh) for d in steps]) if steps else np.zeros((0,) + sh)
    out["solution"] = np.stack([d["solution"] for d in steps]) if steps else np.zeros((0, shape["n"]))
    out["dual"] = np.stack([d["dual"] for d in steps]) if steps else np.zeros((0, shape["m"]))
    out["exit_code"] = np.array([d["exit_code"] for d in steps], np.int32)
    out["iterations"] = np.array([d["iterations"] for d in steps], np.int32)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("dump")
    ap.add_argument("npz")
    ap.add_argument("--preset", required=True)
    ap.add_argument("--adaptive-rho-interval", type=int, default=0,
                    help="interval OSQP used (0 = unknown: the replay tries 25/50/75/100)")
    a = ap.parse_args()
    shape, init, steps = read_dump(a.dump)
    np.savez_compressed(a.npz, **to_npz(shape, init, steps, a.preset, a.adaptive_rho_interval))
    print(f"{a.npz}: Init + {len(steps)} control steps, shape {shape}")


if __name__ == "__main__":
    main()
