"""The C-ABI boundary without a GPU: the library loads, exports every symbol the header
declares, validates configs, and refuses loudly (no CPU fallback) when there is no device."""
import ctypes as C
import os
import re

import pytest

from gs_marl_b200 import abi
from tests._util import make_cfg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "gsmarl_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gsm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = abi.load_library()
    names = header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(abi.SYMBOLS) == names          # python table == header
    assert lib.gsm_abi_version() == abi.GSM_ABI_VERSION


def test_struct_layout_matches_header():
    assert C.sizeof(abi.GsmConfig) == 224
    assert C.sizeof(abi.GsmStepIO) == 80
    assert abi.GsmConfig.dt.offset == 56 and abi.GsmConfig.discrete_u.offset == 160


def _create(cfg, n_envs=4):
    lib = abi.load_library()
    c, keep = cfg.to_c()
    h = C.c_void_p()
    st = lib.gsm_create(C.byref(c), n_envs, 0, 0, C.byref(h))
    msg = lib.gsm_last_error(None).decode()
    if st == 0:
        lib.gsm_destroy(h)
    return st, msg


def test_invalid_configs_are_rejected_before_touching_a_device():
    good = make_cfg("navigation", 3, "f32")
    for kw, frag in [({"max_nbrs": 0}, "max_nbrs"), ({"max_nbrs": 99}, "max_nbrs"),
                     ({"dt": 0.0}, "dt"), ({"contact_margin": 0.0}, "contact_margin"),
                     ({"episode_length": 0}, "episode_length"),
                     ({"mass": [1.0, 0.0, 1.0]}, "mass"),
                     ({"type": [0, 0, 1, 1, 1, 1, 2, 2, 2]}, "agents must be")]:
        st, msg = _create(good.replace(**kw))
        assert st == -1 and frag in msg, (kw, st, msg)
    st, msg = _create(make_cfg("polygon", 4, "f32").replace(n_landmarks=2, size=[0.1] * 6, collide=[1] * 6,
                                                            type=[0, 0, 0, 0, 3, 3], max_nbrs=3))
    assert st == -1 and "polygon" in msg


def test_abi_version_mismatch():
    lib = abi.load_library()
    c, keep = make_cfg("navigation", 3, "f32").to_c()
    c.abi_version = 1
    h = C.c_void_p()
    assert lib.gsm_create(C.byref(c), 4, 0, 0, C.byref(h)) == -3
    c.abi_version = abi.GSM_ABI_VERSION
    c.struct_size = 8
    assert lib.gsm_create(C.byref(c), 4, 0, 0, C.byref(h)) == -3


def test_no_device_means_error_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    st, msg = _create(make_cfg("navigation", 3, "f32"))
    assert st == -5 and "no CPU fallback" in msg
    from gs_marl_b200.environment import MultiAgentGraphConstrainEnv
    with pytest.raises(abi.GsmError):
        MultiAgentGraphConstrainEnv(make_cfg("navigation", 3, "f32"), 4)
    from gs_marl_b200.env_wrappers import GraphVecEnv
    with pytest.raises(abi.GsmError):
        GraphVecEnv(make_cfg("navigation", 3, "f32"), 4)


def test_missing_library_raises(tmp_path):
    with pytest.raises(abi.GsmError):
        abi.load_library(str(tmp_path / "nope.so"))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gs_marl_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inl")):
                s = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", s, flags=re.M), f
                assert "gsm_oracle" not in s, f


def test_ctypes_mirror_matches_the_header_as_compiled_by_gcc(tmp_path):
    """sizeof / offsetof of every struct, taken from include/gsmarl_b200.h by a C compiler, against
    the ctypes mirror in gs_marl_b200/abi.py (a silent mismatch would corrupt weights or pointers)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    checks = [("gsm_config", abi.GsmConfig, ["dtype", "dt", "spawn_extent", "discrete_u", "slot_table"]),
              ("gsm_step_io", abi.GsmStepIO, ["actions", "nbr_feat", "assign"]),
              ("gsm_io_sizes", abi.GsmIoSizes, ["agent_state", "adj_words", "real_bytes"]),
              ("gsm_policy_weights", abi.GsmPolicyWeights, ["n_actions", "ego_w", "nbr_b", "att_b", "head_w",
                                                            "head_b", "value_w", "value_b"]),
              ("gsm_policy_io", abi.GsmPolicyIO, ["logits", "values", "n_rows", "seed", "max_nbrs", "greedy"])]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "gsmarl_b200.h"', 'int main(void) {']
    for cname, _, fields in checks:
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for f in fields:
            lines.append(f'  printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));')
    lines += ['  printf("abi %d\\n", GSM_ABI_VERSION);', '  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                   check=True)
    got = dict(ln.split() for ln in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    assert int(got["abi"]) == abi.GSM_ABI_VERSION
    for cname, ct, fields in checks:
        assert int(got[cname]) == C.sizeof(ct), cname
        for f in fields:
            assert int(got[f"{cname}.{f}"]) == getattr(ct, f).offset, (cname, f)


def test_oracle_types_agree_with_the_public_header(tmp_path):
    """oracle/orc_types.h is declared independently of include/gsmarl_b200.h (the oracle must not
    inherit a mistake of the product's header); a C compiler checks that every field of both config
    and io structs sits at the same offset, and that the oracle's ctypes mirror matches its header."""
    import shutil
    import subprocess
    from oracle import worlds
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    cfg_fields = [n for n, _ in worlds.OrcConfig._fields_]
    io_fields = list(worlds.OrcStepIO.FIELDS)
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "gsmarl_b200.h"', '#include "orc_types.h"',
             'int main(void) {', '  printf("sizeof_cfg %zu %zu\\n", sizeof(gsm_config), sizeof(orc_config));',
             '  printf("sizeof_io %zu %zu\\n", sizeof(gsm_step_io), sizeof(orc_step_io));']
    for f in cfg_fields:
        lines.append(f'  printf("cfg.{f} %zu %zu\\n", offsetof(gsm_config, {f}), offsetof(orc_config, {f}));')
    for f in io_fields:
        lines.append(f'  printf("io.{f} %zu %zu\\n", offsetof(gsm_step_io, {f}), offsetof(orc_step_io, {f}));')
    lines += ['  printf("dims %d %d\\n", GSM_OBS_DIM == ORC_OBS_DIM && GSM_NBR_FEAT_DIM == ORC_NBR_FEAT_DIM && '
              'GSM_MAX_LSA_N == ORC_MAX_LSA_N, (int)GSM_ENT_OBSTACLE == (int)ORC_ENT_OBSTACLE && '
              '(int)GSM_SCN_LINE == (int)ORC_SCN_LINE && (int)GSM_ACT_CONTINUOUS == (int)ORC_ACT_CONTINUOUS);',
              '  return 0;', '}']
    src = tmp_path / "layout2.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout2"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-I",
                    os.path.join(ROOT, "oracle"), str(src), "-o", str(exe)], check=True)
    rows = [ln.split() for ln in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines()]
    for name, a, b in rows:
        if name == "dims":
            assert a == "1" and b == "1"
            continue
        assert a == b, name
        if name.startswith("cfg."):
            assert int(b) == getattr(worlds.OrcConfig, name[4:]).offset, name
        elif name.startswith("io."):
            assert int(b) == getattr(worlds.OrcStepIO, name[3:]).offset, name
    assert dict((r[0], r[2]) for r in rows)["sizeof_cfg"] == str(C.sizeof(worlds.OrcConfig))
    assert (worlds.OBS_DIM, worlds.NBR_FEAT_DIM) == (abi.GSM_OBS_DIM, abi.GSM_NBR_FEAT_DIM)


def test_oracle_never_imports_product():
    for f in os.listdir(os.path.join(ROOT, "oracle")):
        if f.endswith((".py", ".c", ".h")):
            s = open(os.path.join(ROOT, "oracle", f)).read()
            assert not re.search(r"^\s*(from|import)\s+gs_marl_b200\b", s, flags=re.M), f
            assert "include/gsmarl_b200.h" not in s.replace("independent of include/gsmarl_b200.h", ""), f
