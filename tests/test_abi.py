"""CPU-side checks of the drop-in boundary: the shared library builds/loads without a GPU and exports exactly the
symbols that include/geoformer_b200.h declares; the ctypes table mirrors the header; the Python surface keeps the
reference's module paths, config keys and checkpoint schema; the product has no CPU fallback."""
import copy
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "geoformer_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from geoformer_b200 import _lib, build
    build.build()
    lib = _lib.load()                      # dlopen + bind; no compute call (no GPU here)
    declared = _header_functions()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == declared, set(_lib.SIGNATURES) ^ set(declared)
    assert lib.gf_abi_version() == 1
    assert lib.gf_linattn_partial_floats(2, 4800, 8, 32) == 2 * 8 * 38 * (32 * 32 + 32)     # 128-token chunks


def test_header_argument_counts_match_ctypes_table():
    from geoformer_b200 import _lib
    src = open(os.path.join(ROOT, "include", "geoformer_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for name, (_, args) in _lib.SIGNATURES.items():
        m = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % name, src, flags=re.S)
        assert m, name
        params = [p for p in m.group(1).split(",") if p.strip() and p.strip() != "void"]
        assert len(params) == len(args), (name, len(params), len(args))


def test_python_surface_matches_reference():
    from geoformer_b200.model.full_model import GeoFormer
    from geoformer_b200.model.geo_config import default_cfg as geo_cfg
    from geoformer_b200.model.loftr_src.loftr.utils.cvpr_ds_config import default_cfg
    assert set(geo_cfg) == {"layer_names", "nhead", "coarse_thr", "fine_temperature", "fine_thr", "window_size", "topk"}
    assert geo_cfg["coarse_thr"] == 0.2 and geo_cfg["nhead"] == 4 and geo_cfg["layer_names"] == ["self", "cross"] * 2
    assert default_cfg["coarse"]["layer_names"] == ["self", "cross"] * 4 and default_cfg["match_coarse"]["border_rm"] == 2
    conf = copy.deepcopy(default_cfg)
    g = dict(geo_cfg); g["coarse_thr"] = 0.33
    m = GeoFormer(conf, g)
    assert conf["match_coarse"]["thr"] == 0.33                     # full_model.py:31 writes the threshold through
    sd = m.state_dict()
    assert len(sd) == 253
    # wrapper-style load: 'matcher.' prefix, strict=False (geoformer.py:27-30)
    res = m.load_state_dict({"matcher." + k: v.clone() for k, v in sd.items()}, strict=False)
    assert not res.missing_keys and not res.unexpected_keys


def test_no_cpu_fallback():
    from geoformer_b200 import ops
    from geoformer_b200._lib import GeoFormerLibError
    from geoformer_b200.model.full_model import GeoFormer
    from geoformer_b200.model.loftr_src.loftr.utils.cvpr_ds_config import default_cfg
    m = GeoFormer(copy.deepcopy(default_cfg)).eval()
    with pytest.raises(RuntimeError):
        m({"image0": torch.zeros(1, 1, 64, 64), "image1": torch.zeros(1, 1, 64, 64)})
    with pytest.raises((GeoFormerLibError, AssertionError)):
        ops.similarity(torch.zeros(1, 8, 256), torch.zeros(1, 8, 256), 0.1)


def test_training_only_keys_are_refused_not_ignored():
    """scale0 / scale1 / dataset_name (training collations) would change the result if ignored; a lone mask0 is an error
    as in the reference (full_model.py:82-83 reads both).  Padding masks themselves are supported (test_host_forward_
    emulated.py on CPU, test_zz_gpu_masks.py on the GPU)."""
    from geoformer_b200.model.full_model import GeoFormer
    from geoformer_b200.model.geo_config import default_cfg as geo_cfg
    from geoformer_b200.model.loftr_src.loftr.utils.cvpr_ds_config import default_cfg
    m = GeoFormer(copy.deepcopy(default_cfg), dict(geo_cfg)).eval()
    x = torch.zeros(1, 1, 32, 32)
    for k, v in (("scale0", torch.ones(1, 2)), ("scale1", torch.ones(1, 2)), ("dataset_name", ["MegaDepth"])):
        with pytest.raises(NotImplementedError):
            m({"image0": x, "image1": x, k: v})
    with pytest.raises(ValueError):
        m({"image0": x, "image1": x, "mask0": torch.ones(1, 4, 4, dtype=torch.bool)})


def test_product_does_not_import_oracle():
    import subprocess, sys
    code = ("import sys; sys.path.insert(0, %r); import geoformer_b200.model.full_model, geoformer_b200.pipeline, "
            "geoformer_b200.dist; assert not any(m.startswith('oracle') for m in sys.modules), 'oracle imported'") % ROOT
    subprocess.run([sys.executable, "-c", code], check=True)


def test_header_is_valid_c_and_links_from_a_c_host(tmp_path):
    """The boundary is a C ABI: include/geoformer_b200.h must compile as plain C99 (no C++-isms, no torch types) and a C
    host must link against the shared library and call into it (gf_abi_version needs no GPU)."""
    import shutil
    import subprocess
    from geoformer_b200 import build
    if shutil.which("gcc") is None:
        import pytest
        pytest.skip("gcc not available")
    lib = build.build()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "host.c"
    src.write_text('#include "geoformer_b200.h"\n#include <stdio.h>\n'
                   'int main(void) { printf("%d %s\\n", gf_abi_version(), gf_last_error()); return gf_abi_version() == 1 ? 0 : 1; }\n')
    exe = tmp_path / "host"
    inc = os.path.join(root, "include")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, "-fsyntax-only", str(src)], check=True)
    subprocess.run(["gcc", "-std=c99", "-I", inc, str(src), "-o", str(exe), "-L", os.path.dirname(lib),
                    "-l:" + os.path.basename(lib), "-Wl,-rpath," + os.path.dirname(lib)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert out[0] == "1"


def test_mask_entry_points_validate_arguments_without_a_gpu():
    """gf_mask_rows / gf_mask_fill_sim reject bad shapes and misaligned ranges with GF_ERR_ARG before any launch, and an
    empty row set is a no-op (ctypes marshalling of the two new signatures included)."""
    from geoformer_b200 import _lib
    lib = _lib.load()
    x = torch.zeros(8, 16); h = x.half(); m = torch.ones(8, dtype=torch.uint8)
    assert lib.gf_mask_rows(h.data_ptr(), 2, 8, 16, 1, 4, m.data_ptr(), None) == -1          # 2-byte offset
    assert b"multiples of 4 bytes" in lib.gf_last_error()
    assert lib.gf_mask_rows(x.data_ptr(), 4, 8, 16, 10, 8, m.data_ptr(), None) == -1         # column range beyond the row
    assert lib.gf_mask_rows(x.data_ptr(), 8, 8, 16, 0, 8, m.data_ptr(), None) == -1          # element size
    assert lib.gf_mask_rows(None, 4, 0, 8, 0, 8, None, None) == 0                            # no rows: nothing to launch
    sim = torch.zeros(1, 4, 4)
    assert lib.gf_mask_fill_sim(sim.data_ptr(), 1, 70000, 4, m.data_ptr(), m.data_ptr(), -1e9, None) == -1   # grid.y limit
    assert lib.gf_mask_fill_sim(sim.data_ptr(), 0, 4, 4, m.data_ptr(), m.data_ptr(), -1e9, None) == -1
