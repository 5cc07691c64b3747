"""Replaying the hot path of a VAR pass over the C ABI (include/fpq_b200.h).

Two ways in:
  * :class:`DeviceReplay` -- inputs already resident in HBM; every call of the pass is one kernel
    launch on a caller-chosen stream, with no allocation, so the whole pass can be captured in a
    CUDA graph.
  * :class:`HostPipeline` -- HOST buffers in, HOST buffers out: pinned-memory H2D copy, kernel,
    D2H copy, triple-stream pipelined with two staging slots so copies overlap compute.

No arithmetic happens here; there is no CPU fallback.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import _lib as L
from . import ops
from .var_workload import Call

_DT = {"f32": L.FPQ_F32, "f16": L.FPQ_F16}
_TORCH_DT = {"f32": torch.float32, "f16": torch.float16}

# seed-42 sign vector of the reference's 128-block random Hadamard (rotate_utils/hadamard_utils.py:95-96;
# `torch.manual_seed(42); torch.randint(0, 2, (128,))`, 1 -> +1, 0 -> -1, index 0 first)
SIGN_BITS_SEED42_128 = (
    "0100010001000010111010111111110011101000001111101101010110000000"
    "0110111101011101010100101111111111100111111110101101011010100110"
)


def seed42_sign_bits():
    return ops.pack_sign_bits([1.0 if c == "1" else -1.0 for c in SIGN_BITS_SEED42_128])


def gpu_numa_node(device_index: int) -> Optional[int]:
    """NUMA node of the GPU's PCIe function (sysfs), or None when the kernel does not say (-1, single-node hosts)."""
    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        with open(path) as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:  # noqa: BLE001  (no sysfs, no such attribute: nothing to bind to)
        return None


def bind_to_gpu_numa(device_index: int) -> Optional[str]:
    """Pin the calling process to the CPUs of the GPU's NUMA node, so that pinned host buffers allocated AFTERWARDS are
    first-touched on that node and the copy threads run next to it.  With one process per GPU (bench.py under torchrun) this
    spreads the host side of the HostPipeline over the sockets instead of leaving all ranks on the launcher's node.
    Returns the cpulist it bound to, or None when there is nothing to bind to."""
    import os
    node = gpu_numa_node(device_index)
    if node is None:
        return None
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpulist = f.read().strip()
        cpus = set()
        for part in cpulist.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0) or cpus          # stay inside the cgroup's set
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpulist
    except Exception:  # noqa: BLE001
        return None


class DeviceReplay:
    """Launches :class:`Call` s on device pointers.  ``smooth``: dict site -> fp32 CUDA tensor [C]
    (the GALT factor of that site) or None.  ``modulate``: dict site -> (gain, shift), two fp32 CUDA tensors
    [max batches, C] holding ``scale + 1`` and ``shift`` of the adaLN modulate (basic_var.py:258: what ada_lin produced
    for this block), read by "mod_rotate_quant" calls.  One instance serves one device; launches may go to any of its
    streams (the sign-split workspace is kept per stream)."""

    def __init__(self, device: torch.device, smooth: Optional[dict] = None, sign_bits=None, global_clip: bool = True,
                 modulate: Optional[dict] = None):
        if device.type != "cuda":
            raise L.FpqError("DeviceReplay needs a CUDA device (fpqvar_b200 has no CPU fallback)")
        self.device = device
        self.lib = L.lib()
        self.smooth = smooth or {}
        self.modulate = modulate or {}
        self.sign_bits = sign_bits if sign_bits is not None else seed42_sign_bits()
        self.global_clip = global_clip
        # {flag, ticket} of the sign-split whole-tensor clip (fpq_fake_quant_signsplit, FPQ_FLAG_GLOBAL_CLIP), one per stream:
        # two launches on different streams must not share a ticket
        self._flags: dict = {}

    def _flag(self, stream: int) -> int:
        ws = self._flags.get(stream)
        if ws is None:
            ws = self._flags[stream] = torch.zeros(2, dtype=torch.int32, device=self.device)
        return ws.data_ptr()

    def launch(self, call: Call, in_ptr: int, out_ptr: int, stream: int) -> None:
        if torch.cuda.current_device() != self.device.index:
            with torch.cuda.device(self.device):          # the CUDA runtime launches on the calling thread's current device
                return self._launch(call, in_ptr, out_ptr, stream)
        return self._launch(call, in_ptr, out_ptr, stream)

    def _launch(self, call: Call, in_ptr: int, out_ptr: int, stream: int) -> None:
        lib = self.lib
        if call.op == "group":
            n_groups = call.elems // 128
            rc = lib.fpq_fake_quant(in_ptr, out_ptr, n_groups, 128, _DT[call.in_dtype], _DT[call.out_dtype], L.FMT[call.fmt],
                                    L.TIE["kernel"], 0, stream)
            L.check(rc, "fpq_fake_quant")
        elif call.op == "signsplit":
            n_groups = call.elems // 128
            rc = lib.fpq_fake_quant_signsplit(in_ptr, out_ptr, n_groups, 128, _DT[call.in_dtype], _DT[call.out_dtype],
                                              L.SPLIT[call.fmt], L.TIE["kernel"], L.FLAG_GLOBAL_CLIP if self.global_clip else 0,
                                              self._flag(stream) if self.global_clip else None, stream)
            L.check(rc, "fpq_fake_quant_signsplit")
        elif call.op == "rotate_quant":
            s = self.smooth.get(call.site)
            rc = lib.fpq_transform_rotate_quant(in_ptr, s.data_ptr() if s is not None else None, self.sign_bits, out_ptr, None,
                                                call.rows, call.cols, L.FMT[call.fmt], stream)
            L.check(rc, "fpq_transform_rotate_quant")
        elif call.op == "mod_rotate_quant":
            s = self.smooth.get(call.site)
            gain, shift = self.modulate[call.site]
            if gain.shape[0] * call.rows_per_batch < call.rows or gain.shape[1] != call.cols:
                raise L.FpqError(f"modulate tensors of {call.site} are too small for {call.rows} rows / {call.rows_per_batch} per batch")
            rc = lib.fpq_modulate_transform_rotate_quant(in_ptr, gain.data_ptr(), shift.data_ptr(), call.rows_per_batch,
                                                         s.data_ptr() if s is not None else None, self.sign_bits, out_ptr, None,
                                                         call.rows, call.cols, L.FMT[call.fmt], L.MOD_GAIN, stream)
            L.check(rc, "fpq_modulate_transform_rotate_quant")
        else:
            raise L.FpqError(f"unknown op {call.op!r}")


class HostPipeline:
    """HOST-buffer entry of the hot path: for every call, H2D copy of its input from pinned host
    memory, the kernel, and the D2H copy of its output, pipelined over three streams with
    ``slots`` device staging buffers."""

    def __init__(self, device: torch.device, max_in_bytes: int, max_out_bytes: int, smooth: Optional[dict] = None,
                 slots: int = 2, global_clip: bool = True, modulate: Optional[dict] = None):
        self.device = device
        self.replay = DeviceReplay(device, smooth, global_clip=global_clip, modulate=modulate)
        self.slots = slots
        self.d_in = [torch.empty(max_in_bytes, dtype=torch.uint8, device=device) for _ in range(slots)]
        self.d_out = [torch.empty(max_out_bytes, dtype=torch.uint8, device=device) for _ in range(slots)]
        self.s_h2d = torch.cuda.Stream(device)
        self.s_comp = torch.cuda.Stream(device)
        self.s_d2h = torch.cuda.Stream(device)
        self.ev_in = [torch.cuda.Event() for _ in range(slots)]      # input landed
        self.ev_done = [torch.cuda.Event() for _ in range(slots)]    # kernel finished
        self.ev_out = [torch.cuda.Event() for _ in range(slots)]     # output left
        self._used = [False] * slots

    def run(self, calls: Sequence[Call], h_in: Sequence[torch.Tensor], h_out: Sequence[torch.Tensor]) -> None:
        """h_in[i] / h_out[i]: pinned uint8 host tensors holding call i's input / receiving its output."""
        n = self.slots
        for i, call in enumerate(calls):
            k = i % n
            src, dst = h_in[i], h_out[i]
            if not (src.is_pinned() and dst.is_pinned()):
                raise L.FpqError("HostPipeline: host buffers must be pinned")
            din = self.d_in[k][:call.in_bytes]
            dout = self.d_out[k][:call.out_bytes]
            with torch.cuda.stream(self.s_h2d):
                if self._used[k]:
                    self.s_h2d.wait_event(self.ev_done[k])           # staging input free once its kernel ran
                din.copy_(src[:call.in_bytes], non_blocking=True)
                self.ev_in[k].record(self.s_h2d)
            with torch.cuda.stream(self.s_comp):
                self.s_comp.wait_event(self.ev_in[k])
                if self._used[k]:
                    self.s_comp.wait_event(self.ev_out[k])           # staging output free once it was copied out
                self.replay.launch(call, din.data_ptr(), dout.data_ptr(), self.s_comp.cuda_stream)
                self.ev_done[k].record(self.s_comp)
            with torch.cuda.stream(self.s_d2h):
                self.s_d2h.wait_event(self.ev_done[k])
                dst[:call.out_bytes].copy_(dout, non_blocking=True)
                self.ev_out[k].record(self.s_d2h)
            self._used[k] = True

    def synchronize(self) -> None:
        self.s_h2d.synchronize()
        self.s_comp.synchronize()
        self.s_d2h.synchronize()


def run_call(call: Call, x: torch.Tensor, smooth: Optional[torch.Tensor] = None, global_clip: bool = True) -> torch.Tensor:
    """Allocate-and-run convenience used by tests: returns the quantized tensor of ``call``'s out dtype."""
    if call.op == "group":
        return ops.fake_quant(x, call.fmt, 128, "kernel", out_dtype=_TORCH_DT[call.out_dtype])
    if call.op == "signsplit":
        return ops.fake_quant_signsplit(x, call.fmt, 128, "kernel", global_clip=global_clip)
    if call.op == "rotate_quant":
        return ops.transform_rotate_quant(x, smooth, seed42_sign_bits(), call.fmt)
    if call.op == "mod_rotate_quant":
        raise L.FpqError("run_call: a mod_rotate_quant call needs its adaLN tensors; use ops.modulate_transform_rotate_quant")
    raise L.FpqError(f"unknown op {call.op!r}")
