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


class _CpuModel:
    """stands in for imageProcess.EngineModel on a box without a GPU: the state dict, evaluated by the oracle"""
    def __init__(self, sd):
        from oracle import net as N
        self.sd, self.arch = N.to_numpy_state(sd), None


def test_genProcess_drives_the_engine_mirror_end_to_end(monkeypatch):
    """the reference's OWN step builder (procedure.genProcess, procedure.py:156-201) on top of the engine-backed
    runSR / runDN / RGBFilter after install().  This container has no GPU, so the one C-ABI call inside doCrop (moe_run_plan) is
    replaced by the oracle evaluating the same TilePlan on CPU — everything else is the product's host code under the
    reference's closures: getOpt via stepOpts (:171), procSR reading SRopt.ensemble (:67), runSR.sr (:72), RGBFilter (:55), the
    16-bit buffer route toNumPy -> toTorch -> DN -> SR -> toFloat -> toOutput -> BGR2RGB -> toBuffer.  Result: the bytes of the
    stock chain (up to fp32 summation order in the convolutions: at most one 16-bit code value on a few samples)."""
    import importlib
    import types as _t
    import numpy as np
    import torch
    ref = R.load()
    from oracle import net as N, tiling as T
    from moephoto_b200 import install as inst, imageProcess as b_ip
    from moephoto_b200.config import config as b_cfg
    cfg = ref['config']
    cfg.videoPreview, cfg.progressDetail = '', -1
    cfg.crop_sr, cfg.crop_dn = 48, 48
    worker = importlib.import_module('worker')
    worker.context.notifier = _t.SimpleNamespace(send=lambda *a, **k: None)
    procedure = importlib.import_module('procedure')
    frame = np.random.default_rng(5).integers(0, 65536, (64, 96, 3), dtype=np.uint16)
    steps = lambda: [{'op': 'buffer', 'bitDepth': 16}, {'op': 'DN', 'model': 'lite15'}, {'op': 'SR', 'model': 'a', 'scale': 2, 'ensemble': 1}]

    def run_chain():
        process, nodes = procedure.genProcess(steps())
        worker.begin(ref['progress'].Node({}), nodes, -1) if 'progress' in ref else None
        return process((frame.tobytes(), 64, 96))

    progress = importlib.import_module('progress')
    ref['progress'] = progress
    ref['imageProcess'].toBuffer = lambda bitDepth: (lambda im: im.astype(np.uint16).tobytes())      # ndarray.tostring is gone (harness shim 4)
    procedure.toBuffer = ref['imageProcess'].toBuffer
    procedure.previewFormat = ''                                 # no preview file (config.videoPreview is read at import time)
    stock = run_chain()

    def fake_run_plan(model, x, plan, out=None, rows=None, workspace=None):
        oplan = T.Plan()
        oplan.tiles, oplan.pad_h, oplan.pad_w = plan.tiles, plan.pad_h, plan.pad_w
        oplan.out_h, oplan.out_w, oplan.pad_sc, oplan.scale = plan.out_h, plan.out_w, plan.pad_sc, plan.scale
        y = T.do_crop(lambda a: N.forward(model.sd, a, backend='torch'), x.float().numpy(), oplan, ramp=plan.ramp[:plan.pad_sc])
        return torch.from_numpy(y)
    monkeypatch.setattr(b_ip, 'run_plan', fake_run_plan)
    monkeypatch.setattr(b_ip, 'initModel', lambda opt, weights=None, key=None, f=None, args=[]: _CpuModel(b_ip.getStateDict(weights) if isinstance(weights, str) else weights))
    saved = (ref['runSR'].getOpt, ref['runSR'].sr, ref['runDN'].getOpt, ref['imageProcess'].RGBFilter, procedure.RGBFilter)
    b_cfg.freeMemOverride = int(4e9)
    try:
        fake_cfg = _t.SimpleNamespace(cuda=True, fp16=True, deviceId=0, crop_sr=48, crop_dn=48, crop_dns='auto', ensembleSR=0, maxGraphicMemoryUsage=0)
        assert inst.install(fake_cfg) is True
        procedure.RGBFilter = ref['imageProcess'].RGBFilter          # MoePhoto installs before `procedure` is imported; here it already was
        got = run_chain()
    finally:
        ref['runSR'].getOpt, ref['runSR'].sr, ref['runDN'].getOpt, ref['imageProcess'].RGBFilter, procedure.RGBFilter = saved
        b_cfg.freeMemOverride, b_cfg.crop_sr, b_cfg.crop_dn = None, 'auto', 'auto'
        cfg.crop_sr = cfg.crop_dn = 'auto'
    assert isinstance(got, list) and len(got) == 1 and len(got[0]) == len(stock[0]) == 128 * 192 * 3 * 2
    a, b = np.frombuffer(got[0], np.uint16).astype(np.int64), np.frombuffer(stock[0], np.uint16).astype(np.int64)
    assert np.abs(a - b).max() <= 1 and (a != b).mean() < 0.05          # 2e-5 of fp32 summation noise against a 1.5e-5 quantisation step
