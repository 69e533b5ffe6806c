"""Snapshot formats of the reference (src/utils/tester.py:37-51, src/ptqer.py:383-387):
``state_in_fp.pkl`` (fake-quant weights in fp32 + alpha_*), ``state_in_int8.pkl`` (every
quantizer layer's weight replaced by its uint8 codes, PTQConv.py:125-142) and
``state_in_int8_compress.npz``.  Loadable by the reference's ``restore_fp_weight``."""
import numpy as np
import torch


def save(model, filename: str, compress: bool = False) -> None:
    state = {"state_dict": model.state_dict()}
    print(f"Snapshotting to {filename}")
    if compress:
        state["state_dict"] = {k: v.data.cpu().numpy() for k, v in state["state_dict"].items()}
        np.savez_compressed(filename, state)
    else:
        torch.save(state, filename)
