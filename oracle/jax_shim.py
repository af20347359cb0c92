"""NumPy-backed stand-in for the parts of JAX the reference's hot-path files use (TEST INFRASTRUCTURE ONLY).

JAX / jaxlib / jaxopt / mat73 / matplotlib cannot be installed in the build image (no network), so the reference's
own source text cannot be executed on its real runtime.  This module lets the UNMODIFIED files
``/root/reference/Final_python/{solve_helmholtz,nonlinearcg,fwi_loss_function,fwi_script}.py`` be imported and run:
``install()`` injects module objects named ``jax``, ``jax.numpy``, ``jax.lax``, ``jax.scipy.linalg``,
``jax.experimental.sparse[.linalg]``, ``jaxopt``, ``mat73``, ``matplotlib[.pyplot]`` into ``sys.modules``.
``tests/golden/make_ref_golden.py`` uses it to generate the reference-executed fixtures that pin ``oracle/``.

What is reproduced, because the reference's results depend on it:
  * x64 disabled (JAX default): float64 -> float32, complex128 -> complex64, int64 -> int32 on every array
    creation and every operation result; integer arrays meeting a float operand are computed in float32 (JAX's
    promotion lattice), Python scalars are weakly typed (NumPy >= 2 already behaves that way).
    ``config.update("jax_enable_x64", True)`` switches the demotion off, as in JAX.
  * NumPy-style indexing with integer arrays CLAMPS out-of-bounds indices (after wrapping negatives) instead of
    raising -- ``solve_helmholtz.py:226-239`` relies on it.
  * ``x.at[idx].set(v)`` / ``.add(v)`` are functional (copy) updates.
  * ``lax.cond`` / ``lax.scan`` / ``vmap`` / ``jit`` run eagerly (Python branch, loop, per-row loop, identity).
  * ``jax.experimental.sparse.BCOO((data, indices), shape)``, ``.transpose()``, ``BCSR.from_bcoo`` (indices sorted
    lexicographically by (row, col), data carried along -- ``BCOO.sort_indices`` + ``_bcoo_to_bcsr``).
  * ``jax.pure_callback(f, ShapeDtypeStruct, *args)`` calls ``f`` on plain NumPy arrays and casts the result to the
    declared shape / dtype.
What is NOT reproduced: XLA's instruction selection (fused multiply-adds, its own transcendental and reduction
orders), so float32 results agree with real JAX to rounding, not to the bit.  The sparse solve itself is the real
thing: the reference calls SciPy's ``spsolve`` (SuperLU) on the host, and so does this.
"""
from __future__ import annotations

import sys
import types

import numpy as np

_X64 = [False]


def _demote_dtype(dt):
    dt = np.dtype(dt)
    if _X64[0]:
        return dt
    return {np.dtype(np.float64): np.dtype(np.float32), np.dtype(np.complex128): np.dtype(np.complex64),
            np.dtype(np.int64): np.dtype(np.int32), np.dtype(np.uint64): np.dtype(np.uint32)}.get(dt, dt)


def _wrap(x):
    """ndarray / NumPy scalar -> Array with the x64-disabled dtype; everything else unchanged."""
    if isinstance(x, (np.ndarray, np.generic)):
        a = np.asarray(x)
        dt = _demote_dtype(a.dtype)
        if dt != a.dtype:
            a = a.astype(dt)
        return a.view(Array)
    if isinstance(x, tuple):
        return tuple(_wrap(v) for v in x)
    if isinstance(x, list):
        return [_wrap(v) for v in x]
    return x


def _unwrap(x):
    if isinstance(x, Array):
        return x.view(np.ndarray)
    if isinstance(x, (list, tuple)):
        return type(x)(_unwrap(v) for v in x)
    if isinstance(x, dict):
        return {k: _unwrap(v) for k, v in x.items()}
    return x


def _is_float_like(v):
    if isinstance(v, (float, complex)) and not isinstance(v, bool):
        return True
    if isinstance(v, (np.ndarray, np.generic)):
        return np.asarray(v).dtype.kind in "fc"
    return False


def _lattice(args):
    """JAX promotes an integer/bool ARRAY meeting any float operand to float32 (x64 disabled), where NumPy would go
    to float64 and round afterwards: convert such arrays up front so the arithmetic itself runs in float32."""
    if _X64[0] or not any(_is_float_like(a) for a in args):
        return args
    out = []
    for a in args:
        if isinstance(a, (np.ndarray, np.generic)) and np.asarray(a).dtype.kind in "iub":
            a = np.asarray(a).astype(np.float32)
        out.append(a)
    return out


_INT_TO_FLOAT_UFUNCS = {np.true_divide, np.sqrt, np.cos, np.sin, np.exp, np.log}


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIdx(self.arr, idx)


class _AtIdx:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, _unwrap(idx)

    def _apply(self, values, op):
        out = np.array(self.arr.view(np.ndarray), copy=True)
        v = np.asarray(_unwrap(values))
        if op == "set":
            out[self.idx] = v.astype(out.dtype) if v.dtype != out.dtype else v
        else:
            np.add.at(out, self.idx, v.astype(out.dtype))
        return _wrap(out)

    def set(self, values):
        return self._apply(values, "set")

    def add(self, values):
        return self._apply(values, "add")


class Array(np.ndarray):
    """ndarray with JAX's dtype demotion, clamped integer-array gathers and ``.at[]``."""

    __array_priority__ = 100

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        args = [_unwrap(a) for a in inputs]
        if method == "__call__":
            if ufunc in _INT_TO_FLOAT_UFUNCS and not _X64[0]:
                args = [np.asarray(a).astype(np.float32) if isinstance(a, (np.ndarray, np.generic))
                        and np.asarray(a).dtype.kind in "iub" else a for a in args]
            args = _lattice(args)
        if out is not None:
            kwargs["out"] = tuple(_unwrap(o) for o in out)
        res = getattr(ufunc, method)(*args, **kwargs)
        return _wrap(res)

    def __array_function__(self, func, types_, args, kwargs):
        res = func(*_unwrap(list(args)), **_unwrap(kwargs))
        return _wrap(res)

    def __getitem__(self, idx):
        base = self.view(np.ndarray)
        return _wrap(base[_clamp_index(idx, base.shape)])

    def __setitem__(self, idx, v):  # JAX arrays are immutable
        raise TypeError("JAX arrays are immutable; use x.at[idx].set(v)")

    def __iter__(self):
        base = self.view(np.ndarray)
        for i in range(base.shape[0]):
            yield _wrap(base[i])

    @property
    def at(self):
        return _At(self)

    def astype(self, dtype, *a, **k):
        return _wrap(self.view(np.ndarray).astype(_demote_dtype(dtype), *a, **k))

    def __bool__(self):
        return bool(self.view(np.ndarray))

    def __index__(self):
        return self.view(np.ndarray).__index__()

    def __hash__(self):
        return id(self)


def _clamp_index(idx, shape):
    """Out-of-bounds entries of integer index ARRAYS are wrapped (negative) then clamped, as JAX's gather does."""
    tup = idx if isinstance(idx, tuple) else (idx,)
    tup = tuple(_unwrap(t) for t in tup)
    n_consumed = 0
    for t in tup:
        if t is None or t is Ellipsis:
            continue
        if isinstance(t, np.ndarray) and t.dtype == bool:
            n_consumed += t.ndim
        else:
            n_consumed += 1
    out, axis = [], 0
    for t in tup:
        if t is None:
            out.append(t)
        elif t is Ellipsis:
            out.append(t)
            axis += len(shape) - n_consumed
        elif isinstance(t, np.ndarray) and t.dtype == bool:
            out.append(t)
            axis += t.ndim
        elif isinstance(t, (np.ndarray, list)) and np.asarray(t).dtype.kind in "iu":
            a = np.asarray(t)
            dim = shape[axis]
            a = np.where(a < 0, a + dim, a)
            out.append(np.clip(a, 0, dim - 1))
            axis += 1
        else:
            out.append(t)
            axis += 1
    return tuple(out) if isinstance(idx, tuple) else out[0]


# ------------------------------------------------------------------------------------------------ jax.numpy
def _np_wrapper(name):
    fn = getattr(np, name)

    def call(*args, **kwargs):
        args = _unwrap(list(args))
        kwargs = _unwrap(kwargs)
        if "dtype" in kwargs and kwargs["dtype"] is not None:
            kwargs["dtype"] = _demote_dtype(kwargs["dtype"])
        return _wrap(fn(*args, **kwargs))

    call.__name__ = name
    return call


def _default_float():
    return np.float64 if _X64[0] else np.float32


def _creation(name):
    fn = getattr(np, name)

    def call(*args, dtype=None, **kwargs):
        args = _unwrap(list(args))
        if dtype is None:
            r = fn(*args, **_unwrap(kwargs))
            return _wrap(r.astype(_default_float()) if r.dtype == np.float64 else r)
        return _wrap(fn(*args, dtype=_demote_dtype(dtype), **_unwrap(kwargs)))

    return call


def _array(obj, dtype=None, copy=True):
    a = np.array(_unwrap(obj), dtype=None if dtype is None else _demote_dtype(dtype))
    return _wrap(a)


def _linspace(start, stop, num=50, endpoint=True, dtype=None):
    # jnp.linspace works in the (demoted) floating type of its arguments
    ft = _default_float() if dtype is None else _demote_dtype(dtype)
    start, stop = np.asarray(_unwrap(start)).astype(ft), np.asarray(_unwrap(stop)).astype(ft)
    return _wrap(np.linspace(start, stop, num, endpoint=endpoint, dtype=ft))


def _arange(*args, dtype=None):
    """jnp.arange with static arguments delegates to ``np.arange(start, stop, step, dtype=<canonical dtype>)``: with a
    float32 dtype NumPy fills ``start + i * float32(second - first)``, which is why the script's grid
    ``jnp.arange(-0.12, 0.12 + 8e-4, 8e-4)`` has h = 7.9999864e-4 and ends at 0.1199996."""
    args = _unwrap(list(args))
    if dtype is None:
        is_float = any(isinstance(a, float) or (isinstance(a, (np.ndarray, np.generic)) and np.asarray(a).dtype.kind == "f")
                       for a in args)
        dtype = _default_float() if is_float else np.int64
    return _wrap(np.arange(*args, dtype=_demote_dtype(dtype)))


def _take(a, indices, axis=None, mode=None):
    return _wrap(np.take(_unwrap(a), _unwrap(indices), axis=axis))  # in-bounds use only (mode="fill" never triggers)


def _vdot(a, b):
    return _wrap(np.vdot(_unwrap(a), _unwrap(b)))


def _make_jnp():
    m = types.ModuleType("jax.numpy")
    for name in ("mean diff meshgrid maximum minimum abs sign min max squeeze cos sin exp sqrt stack repeat concatenate "
                 "ones_like zeros_like reshape conj real imag transpose sum argmin take_along_axis where ravel isfinite "
                 "cumsum prod dot matmul allclose linalg expand_dims broadcast_to clip sort argsort any all "
                 "angle arctan2 floor ceil round log outer tile isnan nonzero").split():
        setattr(m, name, _np_wrapper(name) if name != "linalg" else None)
    for name in ("ones", "zeros", "full", "eye", "empty"):
        setattr(m, name, _creation(name))
    m.array = m.asarray = _array
    m.linspace, m.arange, m.take, m.vdot = _linspace, _arange, _take, _vdot
    m.pi, m.newaxis, m.inf = float(np.pi), None, float("inf")
    m.float32, m.float64, m.complex64, m.complex128 = np.float32, np.float64, np.complex64, np.complex128
    m.int32, m.int64, m.bool_ = np.int32, np.int64, np.bool_
    m.ndarray = Array
    la = types.ModuleType("jax.numpy.linalg")
    la.norm = _np_wrapper_from(np.linalg.norm)
    la.solve = _np_wrapper_from(np.linalg.solve)
    m.linalg = la
    return m


def _np_wrapper_from(fn):
    def call(*args, **kwargs):
        return _wrap(fn(*_unwrap(list(args)), **_unwrap(kwargs)))
    return call


# ------------------------------------------------------------------------------------------------ jax core
class ShapeDtypeStruct:
    def __init__(self, shape, dtype):
        self.shape, self.dtype = tuple(int(s) for s in shape), np.dtype(dtype)


def _pure_callback(fn, result_shape_dtypes, *args, **kwargs):
    """Host callback: NumPy in, NumPy out, result cast to the declared struct (``solve_helmholtz.py:85-93``)."""
    def to_np(a):
        if isinstance(a, (Array, np.ndarray)):
            return np.asarray(_unwrap(a))
        if isinstance(a, (tuple, list)):
            return type(a)(to_np(v) for v in a)
        if isinstance(a, (int, np.integer)):
            return np.int32(a)
        return a
    res = fn(*[to_np(a) for a in args], **{k: to_np(v) for k, v in kwargs.items()})
    res = np.asarray(res)
    if tuple(res.shape) != result_shape_dtypes.shape:
        raise ValueError("pure_callback: result shape %s != declared %s" % (res.shape, result_shape_dtypes.shape))
    return np.asarray(res, dtype=result_shape_dtypes.dtype).view(Array)


def _jit(fn=None, **kw):
    if fn is None:
        return lambda f: f
    return fn


def _vmap(fn, in_axes=0, out_axes=0):
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = next(np.asarray(_unwrap(a)).shape[ax] for a, ax in zip(args, axes) if ax is not None)
        outs = []
        for i in range(n):
            sl = [a if ax is None else _wrap(np.take(_unwrap(a), i, axis=ax)) for a, ax in zip(args, axes)]
            outs.append(_unwrap(fn(*sl)))
        return _wrap(np.stack([np.asarray(o) for o in outs], axis=out_axes))
    return mapped


def _cond(pred, true_fun, false_fun, *operands, operand="__unset__"):
    if operand != "__unset__":
        operands = (operand,)
    return true_fun(*operands) if bool(pred) else false_fun(*operands)


def _scan(f, init, xs, length=None):
    carry, ys = init, []
    n = length if xs is None else len(xs)
    for i in range(n):
        carry, y = f(carry, None if xs is None else xs[i])
        ys.append(y)
    if all(y is None for y in ys):
        return carry, None
    return carry, _wrap(np.stack([np.asarray(_unwrap(y)) for y in ys]))


class _Config:
    def update(self, key, value):
        if key == "jax_enable_x64":
            _X64[0] = bool(value)


# ------------------------------------------------------------------------------------------------ jax.experimental.sparse
class BCOO:
    def __init__(self, args, shape):
        data, indices = args
        self.data, self.indices, self.shape = _wrap(np.asarray(_unwrap(data))), _wrap(np.asarray(_unwrap(indices))), tuple(int(s) for s in shape)
        self.dtype = self.data.dtype

    def transpose(self, axes=None):
        ind = np.asarray(_unwrap(self.indices))[:, ::-1]
        return BCOO((self.data, np.ascontiguousarray(ind)), shape=self.shape[::-1])

    @property
    def T(self):
        return self.transpose()

    def sort_indices(self):
        ind = np.asarray(_unwrap(self.indices))
        order = np.lexsort((ind[:, 1], ind[:, 0]))
        return BCOO((np.asarray(_unwrap(self.data))[order], ind[order]), shape=self.shape)

    def todense(self):
        out = np.zeros(self.shape, dtype=self.dtype)
        ind = np.asarray(_unwrap(self.indices))
        np.add.at(out, (ind[:, 0], ind[:, 1]), np.asarray(_unwrap(self.data)))
        return _wrap(out)


class BCSR:
    def __init__(self, args, shape):
        self.data, self.indices, self.indptr = (_wrap(np.asarray(_unwrap(a))) for a in args)
        self.shape = tuple(shape)
        self.dtype = self.data.dtype

    @classmethod
    def from_bcoo(cls, arr):
        arr = arr.sort_indices()
        ind = np.asarray(_unwrap(arr.indices))
        counts = np.bincount(ind[:, 0], minlength=arr.shape[0])
        indptr = np.zeros(arr.shape[0] + 1, dtype=np.int32)
        indptr[1:] = np.cumsum(counts).astype(np.int32)
        return cls((arr.data, ind[:, 1].astype(np.int32), indptr), shape=arr.shape)


class _Anything:
    """Absorbs any attribute access / call / subscript / unpack (matplotlib stand-in)."""

    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()

    def __getitem__(self, i):
        return _Anything()

    def __iter__(self):
        return iter((_Anything(), _Anything()))


def _unavailable(name):
    def raiser(*a, **k):
        raise NotImplementedError("%s is not available in the NumPy shim" % name)
    return raiser


def install(mat_loader=None):
    """Put the stand-in modules into ``sys.modules`` (idempotent).  ``mat_loader(path)`` backs ``mat73.loadmat``."""
    jax = types.ModuleType("jax")
    jnp = _make_jnp()
    jax.numpy = jnp
    jax.jit, jax.vmap, jax.pure_callback, jax.ShapeDtypeStruct = _jit, _vmap, _pure_callback, ShapeDtypeStruct
    jax.config = _Config()
    jax.Array = Array
    lax = types.ModuleType("jax.lax")
    lax.cond, lax.scan = _cond, _scan
    jax.lax = lax
    dbg = types.ModuleType("jax.debug")
    dbg.print = lambda fmt, *a, **k: print(fmt.format(*[_unwrap(v) for v in a], **{q: _unwrap(v) for q, v in k.items()}))
    jax.debug = dbg
    jscipy = types.ModuleType("jax.scipy")
    jsl = types.ModuleType("jax.scipy.linalg")
    jsl.solve = _np_wrapper_from(np.linalg.solve)
    jscipy.linalg = jsl
    jax.scipy = jscipy
    exp = types.ModuleType("jax.experimental")
    sparse = types.ModuleType("jax.experimental.sparse")
    sparse.BCOO, sparse.BCSR = BCOO, BCSR
    sl = types.ModuleType("jax.experimental.sparse.linalg")
    sl.spsolve = _unavailable("jax.experimental.sparse.linalg.spsolve")  # imported, never called (solve_helmholtz.py:6)
    sparse.linalg = sl
    exp.sparse = sparse
    jax.experimental = exp
    jaxopt = types.ModuleType("jaxopt")
    jaxopt.LBFGS = _unavailable("jaxopt.LBFGS")
    mat73 = types.ModuleType("mat73")
    mat73.loadmat = (lambda path, **k: mat_loader(path)) if mat_loader else _unavailable("mat73.loadmat")
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    plt.__getattr__ = lambda name: _Anything()
    mpl.pyplot = plt
    mods = {"jax": jax, "jax.numpy": jnp, "jax.lax": lax, "jax.debug": dbg, "jax.scipy": jscipy, "jax.scipy.linalg": jsl,
            "jax.experimental": exp, "jax.experimental.sparse": sparse, "jax.experimental.sparse.linalg": sl,
            "jaxopt": jaxopt, "mat73": mat73, "matplotlib": mpl, "matplotlib.pyplot": plt}
    sys.modules.update(mods)
    return jax


def uninstall():
    for k in [k for k in sys.modules if k == "jax" or k.startswith("jax.") or k in ("jaxopt", "mat73", "matplotlib", "matplotlib.pyplot")]:
        mod = sys.modules[k]
        if getattr(mod, "__file__", None) is None:
            del sys.modules[k]
