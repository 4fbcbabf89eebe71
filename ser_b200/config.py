"""Feature-group switches of the fast profile (mirror of ser.config.FeatureFlags,
ser/_internal/config/schema.py:219-227).  Any object with the same five boolean
attributes -- the reference's own dataclass included -- is accepted wherever flags are taken.
"""

from __future__ import annotations

from dataclasses import dataclass

from . import _native

GROUP_ORDER = ("mfcc", "chroma", "mel", "contrast", "tonnetz")
GROUP_DIMS = {"mfcc": 40, "chroma": 12, "mel": 128, "contrast": 7, "tonnetz": 6}
_GROUP_BITS = {
    "mfcc": _native.FLAG_MFCC,
    "chroma": _native.FLAG_CHROMA,
    "mel": _native.FLAG_MEL,
    "contrast": _native.FLAG_CONTRAST,
    "tonnetz": _native.FLAG_TONNETZ,
}


@dataclass(frozen=True)
class FeatureFlags:
    """Which feature groups are extracted; all on by default, like the reference."""

    mfcc: bool = True
    chroma: bool = True
    mel: bool = True
    contrast: bool = True
    tonnetz: bool = True


def flag_bits(flags) -> int:
    """Bit mask the C ABI takes (SERB_FLAG_*), from any FeatureFlags-shaped object."""
    active = flags if flags is not None else FeatureFlags()
    bits = 0
    for name in GROUP_ORDER:
        if bool(getattr(active, name)):
            bits |= _GROUP_BITS[name]
    return bits


def feature_dim(flags) -> int:
    """Length of the vector the enabled groups produce (ser/_internal/repr/handcrafted.py:46-59)."""
    active = flags if flags is not None else FeatureFlags()
    return sum(GROUP_DIMS[name] for name in GROUP_ORDER if bool(getattr(active, name)))
