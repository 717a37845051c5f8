"""Which GPU a plugin process works on (galsim_plugin._device): the launcher's LOCAL_RANK, an explicit ``b2_device``,
else det_num % n_gpus -- GalSim forks its workers per output file.  Host logic only (stand-in config engine)."""
import sys

import pytest

import pooled_config as pc


@pytest.fixture(autouse=True, scope="module")
def _stand_in_engine_only_here():
    yield
    if pc.STUBS in sys.path:
        sys.path.remove(pc.STUBS)
    for name in [m for m in sys.modules if m in ("galsim", "imsim", "imsim_b200.galsim_plugin")
                 or m.startswith("galsim.") or m.startswith("imsim.")]:
        sys.modules.pop(name, None)


def test_device_of_a_plugin_process(monkeypatch):
    plugin, _ = pc.load_plugin()
    monkeypatch.setattr(plugin, "_N_GPUS", 8)
    monkeypatch.delenv("LOCAL_RANK", raising=False)
    assert plugin._device({"det_num": 94}) == 6
    assert plugin._device({"file_num": 3}) == 3
    assert plugin._device({}) == 0
    assert plugin._device({"det_num": 94, "b2_device": 13}) == 5
    monkeypatch.setenv("LOCAL_RANK", "3")
    assert plugin._device({"det_num": 94}) == 3
    assert plugin._device({"det_num": 94, "b2_device": 1}) == 1
    monkeypatch.setenv("LOCAL_RANK", "not a number")
    assert plugin._device({"det_num": 94}) == 6
