"""Border specifications of the reference (src/border.jl:14-560): Pad, Fill, Inner, NoPad, NA."""
from __future__ import annotations

from . import _abi

_STYLES = {"replicate": _abi.REPLICATE, "circular": _abi.CIRCULAR, "symmetric": _abi.SYMMETRIC,
           "reflect": _abi.REFLECT}
valid_borders = ("replicate", "circular", "reflect", "symmetric")


class AbstractBorder:
    pass


class Pad(AbstractBorder):
    """Pad(style) | Pad(style, lo, hi) | Pad(style, (m, n, …)) [both sides] | Pad(lo, hi) [replicate].
    src/border.jl:14-232."""

    def __init__(self, *args):
        style = "replicate"
        if args and isinstance(args[0], str):
            style, args = args[0], args[1:]
        style = style.lstrip(":")
        if style not in _STYLES:
            raise _abi.ArgumentError(f"border style {style} unrecognized")
        self.style = style
        if len(args) == 0:
            self.lo, self.hi = (), ()
        elif len(args) == 1:      # Pad(style, (m, n, …)) or, for vectors, Pad(style, m)  (src/border.jl:43-45)
            both = (int(args[0]),) if isinstance(args[0], (int,)) else tuple(int(v) for v in args[0])
            self.lo, self.hi = both, both
        elif len(args) == 2 and all(isinstance(a, (tuple, list)) for a in args):
            lo, hi = tuple(int(v) for v in args[0]), tuple(int(v) for v in args[1])
            if not lo:
                lo = (0,) * len(hi)
            if not hi:
                hi = (0,) * len(lo)
            self.lo, self.hi = lo, hi
        else:  # Pad(style, m, n, …)
            both = tuple(int(v) for v in args)
            self.lo, self.hi = both, both
        if len(self.lo) != len(self.hi):
            raise _abi.ArgumentError("lo and hi must have the same length")

    def __repr__(self):
        return f"Pad(:{self.style}, {self.lo}, {self.hi})"

    def to_abi(self, ndim):
        if self.lo and len(self.lo) != ndim:
            raise _abi.ArgumentError(f"{self!r} lacks the proper padding sizes for an array with {ndim} dimensions")
        return _abi.make_border(_STYLES[self.style], 0.0, self.lo or None, self.hi or None)


class Fill(AbstractBorder):
    """Fill(value) | Fill(value, lo, hi) | Fill(value, both).  src/border.jl:354-440."""

    def __init__(self, value, lo=(), hi=None):
        self.value = value
        lo = (int(lo),) if isinstance(lo, int) else tuple(int(v) for v in lo)
        hi = lo if hi is None else ((int(hi),) if isinstance(hi, int) else tuple(int(v) for v in hi))
        self.lo, self.hi = lo, hi

    def __repr__(self):
        return f"Fill({self.value}, {self.lo}, {self.hi})"

    def to_abi(self, ndim):
        if self.lo and len(self.lo) != ndim:
            raise _abi.ArgumentError(f"{self!r} lacks the proper padding sizes for an array with {ndim} dimensions")
        return _abi.make_border(_abi.FILL, float(self.value), self.lo or None, self.hi or None)


class Inner(AbstractBorder):
    """Inner() | Inner(lo, hi) | Inner(both).  src/border.jl:442-560."""

    def __init__(self, lo=(), hi=None):
        lo = (int(lo),) if isinstance(lo, int) else tuple(int(v) for v in lo)
        hi = lo if hi is None else ((int(hi),) if isinstance(hi, int) else tuple(int(v) for v in hi))
        self.lo, self.hi = lo, hi

    def __repr__(self):
        return f"Inner({self.lo}, {self.hi})"

    def to_abi(self, ndim):
        return _abi.make_border(_abi.INNER)


class NoPad(AbstractBorder):
    """NoPad(): the caller has already padded; axes(out) / inds select the valid region
    (src/imfilter.jl:256-278)."""

    def __init__(self, border=None):
        self.border = border

    def __repr__(self):
        return "NoPad()"

    def to_abi(self, ndim):
        return _abi.make_border(_abi.NOPAD)


class NA(AbstractBorder):
    """NA(na=isnan): "Not Available" boundary conditions (src/border.jl:389-405): out-of-range and NA-flagged elements do
    not contribute, the result is renormalised by the weight of the elements that did (src/imfilter.jl:282-318).
    The predicate cannot cross the C ABI as a closure; the three predicates the reference exercises are modes:
    NA() / NA("isnan"), NA("!isfinite"), NA("never") (= `x -> false`)."""
    MODES = {"isnan": 0, "!isfinite": 1, "never": 2}

    def __init__(self, na="isnan"):
        if na not in self.MODES:
            raise _abi.NotSupportedError(f"NA({na!r}): only {sorted(self.MODES)} are available on the device")
        self.na = na
        self.mode = self.MODES[na]

    def __repr__(self):
        return f"NA({self.na})"

    def to_abi(self, ndim):
        raise _abi.ArgumentError("NA() is resolved by imfilter itself (two Fill(0) passes and a division), not by the pad")


def borderinstance(border):
    """src/border.jl:147-156."""
    if isinstance(border, AbstractBorder):
        return border
    if isinstance(border, str):
        b = border.lstrip(":")
        if b in valid_borders:
            return Pad(b)
        if b == "inner":
            raise _abi.ArgumentError("specifying Inner as a string is deprecated, use `imfilter(img, kern, Inner())` instead")
        raise _abi.ArgumentError(f"{border} not a recognized border")
    raise _abi.ArgumentError(f"{border!r} not a recognized border")
