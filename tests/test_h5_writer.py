"""include/samurai/b200_h5.hpp (the HDF5 writer behind samurai::save / dump in the drop-in headers), CPU only: a file it writes
is read back by its own reader and by oracle/h5mini.py, the independent pure-Python reader that parses the reference's golden
files -- same superblock / group / object-header / layout versions."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import h5mini  # noqa: E402


def test_writer_roundtrip_and_h5mini(tmp_path):
    exe = tmp_path / "h5_roundtrip"
    subprocess.run(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), "-o", str(exe), os.path.join(ROOT, "tests", "cpp", "h5_roundtrip.cpp")],
                   check=True)
    out = tmp_path / "t.h5"
    r = subprocess.run([str(exe), str(out)], capture_output=True, text=True)
    assert r.returncode == 0 and "roundtrip OK" in r.stdout, r.stdout + r.stderr
    h = h5mini.H5File(str(out))
    assert h.listdir("/") == ["mesh", "n_process", "restart"]
    assert h.listdir("/mesh") == ["connectivity", "fields", "points", "scaling_factor"]
    assert h.listdir("/mesh/fields") == ["level", "u"]
    pts = h.read("/mesh/points")
    assert pts.shape == (6, 3) and pts.dtype == np.float64 and pts[1, 0] == 0.5
    conn = h.read("/mesh/connectivity")
    assert conn.shape == (2, 4) and conn.dtype == np.uint64 and list(conn[1]) == [1, 4, 5, 2]
    assert list(h.read("/mesh/fields/u")) == [1.25, -3.5]
    assert h.read("/restart/intervals").dtype == np.int64
    level, idx, fields = h5mini.read_samurai_mesh(str(out))
    assert list(level) == [1, 1] and idx.tolist() == [[0, 0], [1, 0]] and set(fields) == {"u", "level"}
    # the structures are the ones the reference's files use: superblock v0, 8-byte offsets, symbol-table root group
    raw = open(out, "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and raw[8] == 0 and raw[13] == 8 and raw[14] == 8
    assert int.from_bytes(raw[40:48], "little") == len(raw)  # end-of-file address
