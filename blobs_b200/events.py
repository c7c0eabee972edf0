"""Mirror of the reference's soft-error history (blobs/src/events.rs): a process-global ring of at most 1000 PhysicsEvents.
The library pushes the reference's two messages (rigid_body.rs:266-275, collider.rs:143-158); this module reads them back."""
from dataclasses import dataclass

from . import _abi as A
from ._lib import load


class Severity:  # events.rs:52-60
    Trace, Debug, Info, Warn, Error, Critical = range(6)


@dataclass
class PhysicsEvent:  # events.rs:42-50
    time_data: tuple
    position: tuple
    message: str
    severity: int
    col_handle: int
    rbd_handle: int


def event_history():
    """EventHistory.events, oldest first"""
    lib = load()
    out = []
    ev = A.PhysicsEvent()
    for i in range(lib.blobs_event_history_len()):
        if lib.blobs_event_history_get(i, ev) != 0:
            break
        out.append(PhysicsEvent((ev.real_time, ev.unpaused_time), (ev.position.x, ev.position.y) if ev.has_position else None,
                                ev.message.decode(), ev.severity, ev.col_handle or None, ev.rbd_handle or None))
    return out


def clear_event_history():
    load().blobs_event_history_clear()
