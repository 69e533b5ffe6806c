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


# ---- extension: true sub-byte packing ---------------------------------------------------------
# The reference stores one uint8 per weight whatever the bit width (PTQConv.py:125-142).  For layers with
# <= 16 levels two codes fit a byte, with <= 4 levels four: ``state_in_packed.npz`` holds, per quantizer
# layer, the packed codes + shape + level count + alpha_w (everything else of the state dict unchanged), and
# ``load_packed`` rebuilds exactly the state dict ``save(model after store_int_weight)`` would have written.
def _bits_for(levels: int) -> int:
    return 2 if levels <= 4 else 4 if levels <= 16 else 8


def bits_needed(codes: np.ndarray, levels: int) -> int:
    """Bits per code for a layer: from the level count, widened when the stored codes do not fit.  (The reference
    derives the codes with alpha_w of the LAST ADMM iterate while the weights are the BEST iterate's
    (EfficientQConv.py:155-158), so round((w/alpha_w + 1)/delta) can exceed levels - 1.)"""
    bits = _bits_for(levels)
    mx = int(codes.max()) if codes.size else 0
    while bits < 8 and mx >= (1 << bits):
        bits *= 2
    return bits


def pack_codes(codes: np.ndarray, levels: int, bits: int = 0) -> np.ndarray:
    """uint8 codes -> bytes with 8 // bits codes each (little end first); bits defaults to the level count's."""
    bits = bits or _bits_for(levels)
    flat = np.ascontiguousarray(codes, dtype=np.uint8).reshape(-1)
    if flat.size and int(flat.max()) >= (1 << bits):
        raise ValueError(f"code {int(flat.max())} does not fit {bits} bits")
    per = 8 // bits
    pad = (-flat.size) % per
    if pad:
        flat = np.concatenate([flat, np.zeros(pad, dtype=np.uint8)])
    flat = flat.reshape(-1, per)
    out = np.zeros(flat.shape[0], dtype=np.uint8)
    for j in range(per):
        out |= (flat[:, j] << (bits * j)).astype(np.uint8)
    return out


def unpack_codes(packed: np.ndarray, levels: int, numel: int, bits: int = 0) -> np.ndarray:
    bits = bits or _bits_for(levels)
    per = 8 // bits
    mask = (1 << bits) - 1
    cols = [(packed >> (bits * j)) & mask for j in range(per)]
    return np.stack(cols, axis=1).reshape(-1)[:numel].astype(np.uint8)


def save_packed(model, filename: str) -> None:
    """Call after ``store_int_weight`` (the quantizer layers' weights are uint8 codes)."""
    from .qconv import PTQConv
    sd = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    out = {}
    packed_keys = []
    for name, m in model.named_modules():
        if isinstance(m, PTQConv) and m.weight.dtype == torch.uint8 and m.qlvl_w <= 256:
            key = f"{name}.weight"
            codes = sd.pop(key)
            bits = bits_needed(codes, m.qlvl_w)
            out[f"packed::{key}"] = pack_codes(codes, m.qlvl_w, bits)
            out[f"shape::{key}"] = np.array(codes.shape, dtype=np.int64)
            out[f"levels::{key}"] = np.int64(m.qlvl_w)
            out[f"bits::{key}"] = np.int64(bits)
            packed_keys.append(key)
    out.update({f"raw::{k}": v for k, v in sd.items()})
    print(f"Snapshotting to {filename} ({len(packed_keys)} packed layers)")
    np.savez_compressed(filename, **out)


def load_packed(filename: str) -> dict:
    """-> {"state_dict": {...}} with uint8 codes for the quantizer layers, as ``state_in_int8.pkl`` holds."""
    z = np.load(filename, allow_pickle=False)
    sd = {}
    for k in z.files:
        kind, key = k.split("::", 1)
        if kind in ("shape", "levels", "bits"):
            continue
        if kind == "raw":
            sd[key] = torch.from_numpy(z[k])
        elif kind == "packed":
            shape = tuple(int(t) for t in z[f"shape::{key}"])
            bits = int(z[f"bits::{key}"]) if f"bits::{key}" in z.files else 0
            codes = unpack_codes(z[k], int(z[f"levels::{key}"]), int(np.prod(shape)), bits)
            sd[key] = torch.from_numpy(codes.reshape(shape))
    return {"state_dict": sd}
