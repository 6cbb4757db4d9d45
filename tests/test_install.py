"""moephoto_b200.install: routing of MoePhoto's module-level entry points (CPU-only checks; needs the reference tree)."""
import sys
import types

import pytest

from oracle import refharness as R

pytestmark = pytest.mark.skipif(not R.available(), reason='reference tree not present')


def test_install_is_a_no_op_without_gpu_fp16():
    ref = R.load()
    from moephoto_b200.install import install
    before = (ref['runSR'].getOpt, ref['runSR'].sr, ref['runDN'].getOpt, ref['imageProcess'].RGBFilter)
    cfg = types.SimpleNamespace(cuda=False, fp16=True)
    assert install(cfg) is False
    assert before == (ref['runSR'].getOpt, ref['runSR'].sr, ref['runDN'].getOpt, ref['imageProcess'].RGBFilter)


def test_install_routes_engine_models_and_falls_through_for_the_rest(monkeypatch):
    ref = R.load()
    from moephoto_b200 import install as inst, runSR as b_runSR, runDN as b_runDN
    calls = []
    monkeypatch.setattr(b_runSR, 'getOpt', lambda o: calls.append(('b_sr', o['model'], o['scale'])) or 'ENGINE_OPT')
    monkeypatch.setattr(b_runDN, 'getOpt', lambda o: calls.append(('b_dn', o['model'])) or 'ENGINE_OPT')
    saved = (ref['runSR'].getOpt, ref['runSR'].sr, ref['runDN'].getOpt, ref['imageProcess'].RGBFilter)
    stock = []
    ref['runSR'].getOpt = lambda o: stock.append(('sr', o['model'], o['scale'])) or 'STOCK_OPT'
    ref['runDN'].getOpt = lambda o: stock.append(('dn', o['model'])) or 'STOCK_OPT'
    try:
        cfg = types.SimpleNamespace(cuda=True, fp16=True, deviceId=0, crop_sr='auto', crop_dn='auto', crop_dns='auto', ensembleSR=0,
                                    maxGraphicMemoryUsage=0)
        assert inst.install(cfg) is True
        assert ref['runSR'].getOpt({'model': 'a', 'scale': 4}) == 'ENGINE_OPT'
        assert ref['runSR'].getOpt({'model': 'lite', 'scale': 2}) == 'ENGINE_OPT'      # MoeNet_lite2 is on the engine too
        assert ref['runSR'].getOpt({'model': 'gan', 'scale': 4}) == 'STOCK_OPT'
        assert ref['runDN'].getOpt({'model': 'lite15'}) == 'ENGINE_OPT'
        assert ref['runDN'].getOpt({'model': 'NAFNet_32'}) == 'STOCK_OPT'
        assert calls == [('b_sr', 'a', 4), ('b_sr', 'lite', 2), ('b_dn', 'lite15')] and stock == [('sr', 'gan', 4), ('dn', 'NAFNet_32')]
    finally:
        ref['runSR'].getOpt, ref['runSR'].sr, ref['runDN'].getOpt, ref['imageProcess'].RGBFilter = saved
