"""CPU-side checks of the drop-in boundary: the library builds/loads, exports every symbol that
include/ssnt_tts_c.h declares, and keeps the reference's abort-on-null behaviour.  No compute
calls are made here (there is no GPU in this container)."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ssnt_tts_c.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(?:void|size_t|unsigned|int|const char)\s*\*?\s*(\w+)\s*\(", text)
    return sorted(set(n for n in names if n.startswith(("ssnt_", "tone_latent_"))))


def test_header_declares_the_seven_reference_symbols():
    # ssnt_tts_c/src/lib.rs:10, 86, 118, 220, 244, 267, 346
    want = {"ssnt_tts_beam_search_decode", "ssnt_extract_best_beam_branch",
            "ssnt_tts_v2_beam_search_decode", "ssnt_order_beam_branch",
            "ssnt_upsample_source_indexes", "tone_latent_beam_search_decode",
            "tone_latent_levenshtein_edit_distance"}
    assert want <= set(declared_symbols())


def test_library_exports_every_declared_symbol(product):
    product.build()
    lib = product.lib()
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in ssnt_tts_c.h but not exported"
    for name in product.C_SYMBOLS:
        assert name in declared_symbols(), f"{name} bound in Python but not declared in the header"
    assert product.backend() == "cuda-sm_100a"


def test_workspace_size_queries(product):
    # pure host arithmetic, no GPU needed
    n = product.forward_backward_workspace_bytes(32, 800, 128)
    assert n >= 32 * 801 * 132 * 4 and n % 256 == 0
    assert product.forward_backward_workspace_bytes(0, 10, 10) > 0


def test_null_pointer_aborts_like_the_reference():
    # ssnt_tts_c/src/lib.rs: every pointer is `assert!(!p.is_null())` → panic → abort.
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from conftest import load_product\n"
        "p = load_product(); import ctypes\n"
        "p.lib().tone_latent_levenshtein_edit_distance(None, None, None, None, 1, 1, None)\n"
    ) % (os.path.join(ROOT, "tests"), ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode != 0
    assert "is_null" in r.stderr


def test_shard_range_partitions_the_batch(product):
    for B in (0, 1, 7, 32, 4096):
        for N in (1, 2, 3, 8):
            spans = [product.shard_range(B, r, N) for r in range(N)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(N - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_host_copy_pool_copies_exactly(product):
    """The copy threads behind the host-pointer lattice calls (csrc/host_copy.cu): every byte arrives, for sizes
    around the piece size, unaligned ends, and from several caller threads at once (a busy pool means the caller
    copies by itself)."""
    import ctypes
    import threading

    import numpy as np
    lib = product.lib()
    lib.ssnt_tts_debug_host_copy.restype = ctypes.c_int
    lib.ssnt_tts_debug_host_copy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    rng = np.random.default_rng(5)
    src = rng.integers(0, 256, 9_000_001, dtype=np.uint8)

    def one(n, off):
        dst = np.zeros(n + 64, np.uint8)
        nt = lib.ssnt_tts_debug_host_copy(dst.ctypes.data + 7, src.ctypes.data + off, n)
        assert nt >= 1
        assert np.array_equal(dst[7:7 + n], src[off:off + n])
        assert not dst[:7].any() and not dst[7 + n:].any()

    for n in (1, 4095, 262144, 262145, 524288, 3_000_000, 9_000_000):
        one(n, 1)
    errs = []

    def worker(k):
        try:
            for i in range(20):
                one(1_000_000 + 4099 * k + i, k)
        except Exception as e:  # noqa: BLE001
            errs.append(e)
    th = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs
