"""C-ABI and host-logic tests that need no GPU: the library loads, exports every symbol the
header declares, fails loudly (no fallback) without a device, and the host-side helpers
(table parse, sharding, DLPack validation) behave."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "shdr.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(shdr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    import shdr
    from shdr import _native
    syms = header_symbols()
    assert len(syms) >= 35
    lib = ctypes.CDLL(_native.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/shdr.h but not exported"
        assert s in _native.SIGNATURES, f"{s} has no ctypes signature in _native.py"
    assert set(_native.SIGNATURES) == set(syms)
    assert _native.lib.shdr_version() == 200
    assert shdr.launch_count() >= 0


def test_only_sm100a_code_in_library():
    from shdr import _native
    out = subprocess.run(["cuobjdump", "--list-elf", _native.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_product_never_imports_oracle_or_torch():
    code = ("import sys; sys.path.insert(0, %r); import shdr; "
            "bad=[m for m in sys.modules if m.split('.')[0] in ('oracle','torch','tensorflow','triton')]; "
            "print(bad); sys.exit(1 if bad else 0)" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    for fn in os.listdir(os.path.join(ROOT, "singlehdr-tf2_b200")):
        if fn.endswith(".py"):
            src = open(os.path.join(ROOT, "singlehdr-tf2_b200", fn)).read()
            assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), fn


def _has_gpu():
    import shdr
    return shdr.device_count() > 0


def test_fails_loudly_without_device():
    import shdr
    if _has_gpu():
        pytest.skip("a GPU is present")
    with pytest.raises(shdr.ShdrError):
        shdr.require_gpu()
    with pytest.raises(shdr.ShdrError):
        shdr.DeviceArray.empty((1, 4, 4, 3))
    with pytest.raises(shdr.ShdrError):
        shdr.frontend_host(np.zeros((1, 4, 4, 3), np.float32))
    # host pointers are rejected by the device-pointer ABI: there is no CPU path behind it
    from shdr import _native as N
    a = np.zeros(16, np.float32)
    rc = N.lib.shdr_apply_rf_f32(a.ctypes.data, a.ctypes.data, a.ctypes.data, 1, 16, 16, None)
    assert rc < 0 and N.last_error()


def test_argument_validation_messages():
    from shdr import _native as N
    rc = N.lib.shdr_soft_hist_f32(None, None, 1, 4, 4, 3, 0, 0, 0, 0, None)
    assert rc == N.ERR_INVALID
    rc = N.lib.shdr_frontend_f32(None, None, 1, 1, 4, 0, None)        # h < 2
    assert rc == N.ERR_INVALID
    rc = N.lib.shdr_frontend_f32(None, None, 0, 8, 8, 0, None)        # empty batch is a no-op
    assert rc == N.OK
    rc = N.lib.shdr_apply_rf_f32(None, None, None, 0, 100, 1024, None)
    assert rc == N.OK
    a = np.zeros(4, np.float32)
    rc = N.lib.shdr_increase_f32(a.ctypes.data, a.ctypes.data, 1, 1, None)   # k < 2
    assert rc == N.ERR_INVALID and "k=1" in N.last_error()
    g = np.zeros(10, np.float32)
    rc = N.lib.shdr_set_emor_table(g.ctypes.data, g.ctypes.data, 10, 11)
    assert rc == N.ERR_INVALID
    # the fused conv1 entry points (two bf16 operand images of the 7x7x96x64 kernel: single CTA and CTA pair)
    assert N.lib.shdr_conv1_packed_bytes() == 2 * 49 * 96 * 64 * 2
    rc = N.lib.shdr_frontend_conv1_f32(None, None, None, None, 0, None, 1, 8, 8, None)
    assert rc == N.ERR_INVALID and "NULL" in N.last_error()
    rc = N.lib.shdr_frontend_conv1_f32(None, None, None, None, 0, None, 0, 8, 8, None)   # empty batch is a no-op
    assert rc == N.OK
    rc = N.lib.shdr_conv1_pack_weights_f32(None, None, None)
    assert rc == N.ERR_INVALID


def test_dlpack_rejects_cpu_tensors():
    import torch
    import shdr
    t = torch.zeros(1, 4, 4, 3)
    with pytest.raises(shdr.ShdrError, match="kDLCUDA"):
        shdr.frontend(t)


def test_parse_invemor_product(tmp_path, emor):
    import shdr
    b, g0, hinv = emor
    cols = [("B =", b), ("g0 =", g0)] + [(f"hinv({i + 1})=", hinv[:, i]) for i in range(11)]
    lines = []
    for tag, v in cols:
        lines.append(tag + " ")
        lines += ["   ".join(f"{float(t):.9e}" for t in r) for r in v.reshape(256, 4)]
    p = tmp_path / "invemor.txt"
    p.write_text("\n".join(lines) + "\n")
    b2, g2, h2 = shdr.parse_invemor(str(p))
    assert np.array_equal(b2, b) and np.array_equal(g2, g0) and np.array_equal(h2, hinv)
    assert shdr.parse_invemor(str(p))[2] is h2                         # cached, not re-parsed
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        assert np.array_equal(shdr.AEInvcrfDecodeNet().parse_invemor()[1], g0)   # CWD-relative default
    finally:
        os.chdir(cwd)
    with pytest.raises(FileNotFoundError):
        shdr.parse_invemor(str(tmp_path / "missing.txt"))
    (tmp_path / "bad.txt").write_text("B = \n1 2 3 4\n")
    with pytest.raises(ValueError):
        shdr.parse_invemor(str(tmp_path / "bad.txt"))


def test_shard_range_and_tiles():
    import shdr
    for n in (0, 1, 7, 8, 33):
        for world in (1, 2, 4, 8):
            parts = [shdr.shard_range(n, world, r) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    tiles = shdr.row_tiles(2160, 8, 7, 8)
    assert tiles[0][2] == 0 and tiles[-1][3] == 2160
    assert all(t[2] == max(0, t[0] - 7) and t[3] == min(2160, t[1] + 8) for t in tiles)
    with pytest.raises(ValueError):
        shdr.shard_range(4, 2, 2)


def test_sass_of_the_fused_conv1_uses_tcgen05_and_bulk_copies():
    """The shipped library really drives the 5th-generation tensor cores: the SASS of libshdr.so holds tcgen05 MMAs in
    CTA-pair form (UTCHMMA.2CTA), multicast commits (UTCBAR), tensor-memory loads (LDTM) and allocation (UTCATOMSWS), and
    1-D bulk copies (UBLKCP) -- the mnemonics B200_PROFILING.md lists as the proof of tcgen05 / TMA."""
    import shutil
    import subprocess
    from shdr import _native
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _native.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    for mnemonic in ("UTCHMMA.2CTA", "UTCHMMA ", "UTCBAR.2CTA.MULTICAST", "LDTM.x16", "UTCATOMSWS", "UBLKCP.S.G", "SYNCS.ARRIVE.TRANS64"):
        assert mnemonic in sass, f"{mnemonic} missing from the SASS of libshdr.so"
    assert "sm_100a" in subprocess.run([cuobjdump, "-lelf", _native.LIB_PATH], capture_output=True, text=True).stdout
