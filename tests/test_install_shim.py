"""aimet_b200.install registers the drop-ins under the reference's two native module names."""
import sys
import types


def test_install_registers_reference_module_names(monkeypatch):
    import aimet_b200.install as inst
    if "aimet_common" not in sys.modules:
        monkeypatch.setitem(sys.modules, "aimet_common", types.ModuleType("aimet_common"))
    atq, pymo = inst.install(extra_pymo_names={"QnnDatatype": object})
    try:
        from aimet_common.AimetTensorQuantizer import AimetTensorQuantizer   # noqa: F401
        import aimet_common._libpymo as lp
        names = {}
        exec("from aimet_common._libpymo import *", names)   # what aimet_common/libpymo.py:43 does
        for n in ("TfEncoding", "QuantizationMode", "RoundingMode", "TensorQuantizer", "EncodingAnalyzerForPython",
                  "TensorQuantizationSimForPython", "QnnDatatype"):
            assert n in names, n
        assert lp.QuantizationMode.QUANTIZATION_TF_ENHANCED == 1 and lp.RoundingMode.ROUND_STOCHASTIC == 1
        e = lp.TfEncoding()
        e.min, e.max, e.bw = -1.0, 1.0, 8
        assert (e.min, e.max, e.delta, e.offset, e.bw) == (-1.0, 1.0, 0.0, 0.0, 8)
    finally:
        sys.modules.pop("aimet_common.AimetTensorQuantizer", None)
        sys.modules.pop("aimet_common._libpymo", None)
