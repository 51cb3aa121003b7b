"""The OpSlot / KERNEL_REGISTRY mirror keeps the reference's contract (veomni/ops/dispatch.py, kernel_registry.py)."""
import pytest

from veomni_b200 import registry as R


def test_eager_resolves_to_none_and_unknown_raises_keyerror():
    assert R.KERNEL_REGISTRY.resolve("rms_norm", "standard", "eager") is None
    with pytest.raises(KeyError):
        R.KERNEL_REGISTRY.resolve("rms_norm", "standard", "does_not_exist")
    assert "b200" in R.KERNEL_REGISTRY.list_available("rms_norm", "standard")
    for op, var in (("rotary_pos_emb", "full"), ("swiglu_mlp", "standard"), ("moe_experts", "standard")):
        assert "b200" in R.KERNEL_REGISTRY.list_available(op, var)


def test_unbound_slot_raises_and_eager_binding_is_falsy():
    slot = R.OpSlot("rms_norm", "standard")
    assert not slot.use_non_eager_impl
    with pytest.raises(RuntimeError):
        slot(1, 2, 3)
    slot.bind("eager")
    assert not slot.use_non_eager_impl and slot.bound_kernel() is None


def test_hardware_gate_fails_loudly_without_a_b200():
    import torch

    if torch.cuda.is_available() and torch.cuda.get_device_capability()[0] >= 10:
        pytest.skip("a Blackwell GPU is present")
    slot = R.OpSlot("rms_norm", "standard")
    with pytest.raises(RuntimeError):
        slot.bind("b200")  # no silent CPU fallback


def test_duplicate_registration_rejected():
    reg = R.KernelRegistry()
    spec = R.KernelSpec("x", "op", "v", lambda: (lambda: 1), R.HardwareRequirement("any"))
    reg.register(spec)
    with pytest.raises(ValueError):
        reg.register(spec)
    reg.register(spec, force=True)
    assert reg.resolve("op", "v", "x")() == 1
