"""CPU suite: libaz_b200.so builds, loads and exports every symbol include/az_b200.h declares;
without a GPU every handle constructor fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for fn in sorted(os.listdir(os.path.join(ROOT, "include"))):
        if not fn.startswith("az_b200"):
            continue
        text = open(os.path.join(ROOT, "include", fn)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names += re.findall(r"^\s*AZ_API\s+[a-z_ \*]*?(az_[a-z0-9_]+)\s*\(", text, flags=re.M)
    return sorted(set(names))


@pytest.fixture(scope="module")
def libpath():
    from alphazero_risk_b200 import build
    return build.build()


def test_exports_every_declared_symbol(libpath):
    L = C.CDLL(libpath)
    names = declared_symbols()
    assert len(names) >= 15
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_default_rules_match_reference_settings(libpath):
    from alphazero_risk_b200 import api
    r = api.default_rules()
    assert (r.allow_yield, r.limit_reinforcement, r.limit_attack, r.max_game_rounds, r.min_unit_move) == (1, 1, 0, 58, 3)
    assert (r.mcts_simulations, r.threads_per_mcts, r.temperature_threshold) == (32, 2, 43)
    assert abs(r.cpuct - 1.1) < 1e-7 and abs(r.dir_noise_value - 0.3) < 1e-7 and r.dir_noise_epsi == 0.25


def test_no_cpu_fallback(libpath):
    from alphazero_risk_b200 import api
    if api.lib().az_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(api.AzError):
        api.Env(4)


def test_product_never_imports_oracle():
    """the oracle is test infrastructure: nothing under alphazero_risk_b200/ may reference it"""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "alphazero_risk_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                t = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"risk_oracle|pyoracle|libref_oracle|from oracle|import oracle", t):
                    bad.append(f)
    assert not bad, bad
