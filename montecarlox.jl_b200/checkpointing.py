"""Checkpoint / restore (src/infrastructure/checkpointing.jl:20-111) for device-backed state.

`checkpoint_` serialises a dict of objects with pickle through a tmp file + atomic rename, exactly as
the reference serialises a NamedTuple (:48-56).  Device systems pickle as (host spins, couplings,
Philox seed + next sweep, labels, counters): because the RNG is counter-based, a restored run
continues the trajectory bit for bit (cf. test/test_checkpointing.jl:100-147)."""
import os
import pickle
import tempfile


class CheckpointSession:
    """checkpointing.jl:20-40: fields of the stored state are reachable as attributes."""

    def __init__(self, file, state):
        object.__setattr__(self, "file", file)
        object.__setattr__(self, "_state", dict(state))

    def __getattr__(self, name):
        st = object.__getattribute__(self, "_state")
        if name in st:
            return st[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        self._state[name] = value


def checkpoint_(ckpt, **kwargs):
    """checkpoint!(ckpt; kwargs...) (:48-56)"""
    state = dict(ckpt._state)
    state.update(kwargs)          # merged into the written snapshot only; the session's own state is left alone (:49-50)
    tmp = ckpt.file + ".tmp"
    with open(tmp, "wb") as io:
        pickle.dump(state, io, protocol=pickle.HIGHEST_PROTOCOL)
    os.replace(tmp, ckpt.file)
    return ckpt.file


def init_checkpoint(file=None, state=None, **kwargs):
    """init_checkpoint(file, state; kwargs...) (:70-79)"""
    if file is None:
        file = os.path.join(tempfile.mkdtemp(), "ckpt.mcx")
    d = os.path.dirname(file)
    if d:
        os.makedirs(d, exist_ok=True)
    ckpt = CheckpointSession(file, state or {})
    checkpoint_(ckpt, **kwargs)
    return ckpt


def restore_checkpoint(path):
    """restore_checkpoint(path) (:95-101)"""
    with open(path, "rb") as io:
        state = pickle.load(io)
    return CheckpointSession(path, state)


def finalize_(ckpt):
    """finalize!(ckpt) (:108-111)"""
    try:
        os.remove(ckpt.file)
    except FileNotFoundError:
        pass
    return None
