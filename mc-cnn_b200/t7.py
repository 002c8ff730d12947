"""Reader (and a writer, for fixtures and exports) of Torch7's ASCII serialisation -- the format of the reference's
network files: ``torch.save(fname, obj, 'ascii')`` with obj = {net_te, [net_te2,] opt} (main.lua:587-600), read back by
``torch.load(opt.net_fname, 'ascii')`` (main.lua:893-901).

The format lives in Torch7 itself (torch7/File.lua ``writeObject`` / ``readObject``, torch7/lib/TH/THDiskFile.c, the
tensors' and storages' ``write`` methods), a dependency that is NOT under /root/reference and whose version the reference
does not pin; no trained net can be fetched offline.  This module restates the published format -- **parity unpinned**: it
is checked against hand-written files that follow the layout below and through write -> read round trips, not against a file
produced by Torch7.

  object   :=  TYPE ...                         every scalar is text; each write call ends with '\\n' (auto-spacing), the
                                                values of ONE call are separated by ' '
  0 nil | 1 number: '%.17g' | 5 boolean: 0/1 | 2 string: LEN '\\n' LEN raw bytes
  3 table  :=  INDEX [ COUNT (key object, value object) * COUNT ]      -- the bracket only the first time INDEX is seen
  4 torch  :=  INDEX [ LEN 'V <version>' LEN '<class name>' payload ]  -- idem; files without the 'V ' string are version 0
  payload of torch.*Tensor   :=  NDIM / sizes (one line, absent when NDIM = 0) / strides (idem) / OFFSET (1-based) / storage object
  payload of torch.*Storage  :=  SIZE / SIZE values on one line ('%.9g' float, '%.17g' double, integers as such)
  payload of any other class :=  one table object holding the instance's fields (nn modules have no write method)

Shared references (the tied weights of main.lua:704-722 share storages) are kept: tensors that point at the same storage
come back as numpy views of one array.
"""
import collections
import io
import os
import re

import numpy as np

TYPE_NIL, TYPE_NUMBER, TYPE_STRING, TYPE_TABLE, TYPE_TORCH, TYPE_BOOLEAN = 0, 1, 2, 3, 4, 5
TYPE_FUNCTION, TYPE_LEGACY_RECUR_FUNCTION, TYPE_RECUR_FUNCTION = 6, 7, 8

_ELEM = {"Double": np.float64, "Float": np.float32, "Half": np.float16, "Long": np.int64, "Int": np.int32,
         "Short": np.int16, "Char": np.int8, "Byte": np.uint8,
         "Cuda": np.float32, "CudaDouble": np.float64, "CudaHalf": np.float16, "CudaLong": np.int64, "CudaInt": np.int32,
         "CudaShort": np.int16, "CudaChar": np.int8, "CudaByte": np.uint8}
_TENSOR_RE = re.compile(r"^torch\.(\w*)Tensor$")
_STORAGE_RE = re.compile(r"^torch\.(\w*)Storage$")


class T7Error(ValueError):
    pass


class T7Table(dict):
    """A Lua table.  Integer-valued number keys come back as Python ints (1-based, as in Lua)."""

    def array(self):
        """The values of keys 1..n (ipairs order)."""
        out, i = [], 1
        while i in self:
            out.append(self[i])
            i += 1
        return out


class T7Object:
    """An instance of a torch class that is neither a tensor nor a storage (nn.Sequential, cudnn.SpatialConvolution ...):
    `typename` and the instance's fields."""

    def __init__(self, typename, fields=None, version=1):
        self.typename = typename
        self.fields = fields if fields is not None else T7Table()
        self.version = version

    def __getitem__(self, k):
        return self.fields[k]

    def get(self, k, default=None):
        return self.fields.get(k, default)

    def __repr__(self):
        return "T7Object(%s, %s)" % (self.typename, sorted(map(str, self.fields)))


# --------------------------------------------------------------------------------------------------------------- reading
class _Reader:
    def __init__(self, buf):
        self.b = buf
        self.p = 0
        self.memo = {}

    def _skip_ws(self):
        b, p, n = self.b, self.p, len(self.b)
        while p < n and b[p] in b" \t\r\n":
            p += 1
        self.p = p

    def token(self):
        self._skip_ws()
        b, p, n = self.b, self.p, len(self.b)
        q = p
        while q < n and b[q] not in b" \t\r\n":
            q += 1
        if q == p:
            raise T7Error("t7: unexpected end of file at byte %d" % p)
        self.p = q
        return b[p:q]

    def integer(self):
        t = self.token()
        try:
            return int(t)
        except ValueError:
            raise T7Error("t7: expected an integer at byte %d, found %r" % (self.p - len(t), t[:20])) from None

    def number(self):
        t = self.token()
        try:
            return float(t)                              # also 'nan', 'inf', '-inf' as printf writes them
        except ValueError:
            raise T7Error("t7: expected a number at byte %d, found %r" % (self.p - len(t), t[:20])) from None

    def raw(self, n):
        """The n bytes of a string: they start right after the single '\\n' that ended the length."""
        if self.b[self.p:self.p + 2] == b"\r\n":
            self.p += 2
        elif self.b[self.p:self.p + 1] == b"\n":
            self.p += 1
        if self.p + n > len(self.b):
            raise T7Error("t7: string of %d bytes runs past the end of the file" % n)
        s = self.b[self.p:self.p + n]
        self.p += n
        return s

    def values(self, n, dtype):
        """n numbers written by one call: a few tokens (sizes, strides) or one long line (a storage; wrapped lines tolerated)."""
        integer = np.issubdtype(dtype, np.integer)
        if n <= 16:
            toks = [self.token() for _ in range(n)]
        else:
            toks = []
            while len(toks) < n:
                self._skip_ws()
                e = self.b.find(b"\n", self.p)
                e = len(self.b) if e < 0 else e
                if e == self.p:
                    raise T7Error("t7: %d values expected, %d found" % (n, len(toks)))
                toks += self.b[self.p:e].split()
                self.p = e
            if len(toks) != n:
                raise T7Error("t7: %d values expected on the lines ending at byte %d, %d found" % (n, self.p, len(toks)))
        try:
            if integer:
                return np.array([int(t) for t in toks], dtype=np.int64).astype(dtype)
            return np.array(toks, dtype=np.float64).astype(dtype)
        except (ValueError, OverflowError):
            raise T7Error("t7: malformed number near byte %d" % self.p) from None

    def string(self):
        return self.raw(self.integer()).decode("latin-1")

    def obj(self):
        t = self.integer()
        if t == TYPE_NIL:
            return None
        if t == TYPE_NUMBER:
            v = self.number()
            return int(v) if v == v and abs(v) < 2 ** 53 and v == int(v) else v
        if t == TYPE_BOOLEAN:
            return self.integer() != 0
        if t == TYPE_STRING:
            return self.string()
        if t == TYPE_TABLE:
            idx = self.integer()
            if idx in self.memo:
                return self.memo[idx]
            tab = self.memo[idx] = T7Table()
            for _ in range(self.integer()):
                k = self.obj()
                tab[k] = self.obj()
            return tab
        if t == TYPE_TORCH:
            idx = self.integer()
            if idx in self.memo:
                return self.memo[idx]
            s = self.string()
            m = re.match(r"^V (.*)$", s)
            version, cls = (int(m.group(1)), self.string()) if m else (0, s)
            return self._torch(idx, cls, version)
        if t in (TYPE_FUNCTION, TYPE_LEGACY_RECUR_FUNCTION, TYPE_RECUR_FUNCTION):
            raise T7Error("t7: the file holds a serialised Lua function; not supported")
        raise T7Error("t7: unknown type tag %d at byte %d" % (t, self.p))

    def _torch(self, idx, cls, version):
        m = _STORAGE_RE.match(cls)
        if m and m.group(1) in _ELEM:
            n = self.integer()
            st = self.values(n, _ELEM[m.group(1)]) if n else np.empty(0, _ELEM[m.group(1)])
            self.memo[idx] = st
            return st
        m = _TENSOR_RE.match(cls)
        if m and m.group(1) in _ELEM:
            dt = np.dtype(_ELEM[m.group(1)])
            nd = self.integer()
            size = [int(v) for v in self.values(nd, np.int64)] if nd else []
            stride = [int(v) for v in self.values(nd, np.int64)] if nd else []
            off = self.integer() - 1
            st = self.obj()
            if nd == 0 or st is None:
                a = np.empty((0,), dt)
            else:
                if st.dtype != dt:
                    raise T7Error("t7: %s over a %s storage" % (cls, st.dtype))
                need = off + sum((s - 1) * k for s, k in zip(size, stride)) + 1 if all(size) else 0
                if off < 0 or need > st.size or any(k < 0 for k in stride):
                    raise T7Error("t7: %s of size %s / stride %s / offset %d does not fit its storage of %d" %
                                  (cls, size, stride, off + 1, st.size))
                a = np.lib.stride_tricks.as_strided(st[off:], shape=size, strides=[k * dt.itemsize for k in stride])
            self.memo[idx] = a
            return a
        o = self.memo[idx] = T7Object(cls, None, version)
        f = self.obj()
        if not isinstance(f, T7Table):
            raise T7Error("t7: instance of %s without a field table" % cls)
        o.fields = f
        return o


def loads(buf):
    """The object held by the bytes of an ascii .t7 file."""
    if isinstance(buf, str):
        buf = buf.encode("latin-1")
    r = _Reader(bytes(buf))
    v = r.obj()
    r._skip_ws()
    if r.p != len(r.b):
        raise T7Error("t7: %d trailing bytes after the object" % (len(r.b) - r.p))
    return v


def load(fname):
    """``torch.load(fname, 'ascii')``."""
    with open(fname, "rb") as f:
        head = f.read(16)
        if head and not re.match(rb"^\s*\d+\s", head):
            raise T7Error("t7: %s is not an ASCII-mode Torch7 file (the reference saves its nets with 'ascii', main.lua:598)" % fname)
        return loads(head + f.read())


# --------------------------------------------------------------------------------------------------------------- writing
_TNAME = {np.dtype(np.float64): "Double", np.dtype(np.float32): "Float", np.dtype(np.float16): "Half",
          np.dtype(np.int64): "Long", np.dtype(np.int32): "Int", np.dtype(np.int16): "Short", np.dtype(np.int8): "Char",
          np.dtype(np.uint8): "Byte"}


def _fmt(v, dt):
    if dt == np.float32:
        return "%.9g" % v
    if dt.kind == "f":
        return "%.17g" % v
    return "%d" % v


class _Writer:
    def __init__(self, cuda):
        self.o = io.BytesIO()
        self.memo = {}            # id(python object) / storage root address -> index
        self.keep = []
        self.n = 0
        self.cuda = cuda

    def line(self, *vals):
        self.o.write((" ".join(str(v) for v in vals) + "\n").encode("latin-1"))

    def string(self, s):
        b = s.encode("latin-1")
        self.line(len(b))
        self.o.write(b + b"\n")

    def _index(self, key, tag):
        """Writes TYPE and INDEX; True when the body must follow (first visit)."""
        self.line(tag)
        if key in self.memo:
            self.line(self.memo[key])
            return False
        self.n += 1
        self.memo[key] = self.n
        self.line(self.n)
        return True

    def _cls(self, base, dt):
        name = _TNAME[np.dtype(dt)]
        if self.cuda and name == "Float":
            name = "Cuda"
        return "torch.%s%s" % (name, base)

    def tensor(self, a):
        self.keep.append(a)
        if not self._index(("t", id(a)), TYPE_TORCH):
            return
        self.string("V 1")
        self.string(self._cls("Tensor", a.dtype))
        if a.size == 0:
            self.line(0)
            self.line(1)
            self.line(TYPE_NIL)
            return
        root = a
        while isinstance(root.base, np.ndarray):
            root = root.base
        if not root.flags.c_contiguous or any(s % a.itemsize or s < 0 for s in a.strides):
            root = a = np.ascontiguousarray(a)
            self.keep.append(a)
        off = (a.__array_interface__["data"][0] - root.__array_interface__["data"][0]) // a.itemsize
        self.line(a.ndim)
        self.line(*a.shape)
        self.line(*[s // a.itemsize for s in a.strides])
        self.line(off + 1)
        self.keep.append(root)
        if self._index(("s", root.__array_interface__["data"][0], root.dtype.str), TYPE_TORCH):
            self.string("V 1")
            self.string(self._cls("Storage", root.dtype))
            flat = root.reshape(-1)
            self.line(flat.size)
            self.line(*[_fmt(v, flat.dtype) for v in flat.tolist()])

    def obj(self, v):
        if v is None:
            self.line(TYPE_NIL)
        elif isinstance(v, (bool, np.bool_)):
            self.line(TYPE_BOOLEAN)
            self.line(1 if v else 0)
        elif isinstance(v, (int, float, np.integer, np.floating)):
            self.line(TYPE_NUMBER)
            self.line("%.17g" % v)
        elif isinstance(v, str):
            self.line(TYPE_STRING)
            self.string(v)
        elif isinstance(v, np.ndarray):
            self.tensor(v)
        elif isinstance(v, T7Object):
            self.keep.append(v)
            if self._index(("o", id(v)), TYPE_TORCH):
                self.string("V %d" % v.version)
                self.string(v.typename)
                self.obj(v.fields)
        elif isinstance(v, (dict, list, tuple)):
            self.keep.append(v)
            if self._index(("o", id(v)), TYPE_TABLE):
                items = list(v.items()) if isinstance(v, dict) else list(enumerate(v, 1))
                self.line(len(items))
                for k, x in items:
                    self.obj(k)
                    self.obj(x)
        else:
            raise T7Error("t7: cannot serialise a %s" % type(v).__name__)


def dumps(obj, cuda=True):
    """The ascii serialisation of `obj`: None / bool / numbers / str / dict (list, tuple: keys 1..n) / numpy arrays
    (fp32 arrays as torch.CudaTensor when `cuda`, as the reference's nets hold them, else torch.FloatTensor; views of one
    array share its storage) / :class:`T7Object`."""
    w = _Writer(cuda)
    w.obj(obj)
    return w.o.getvalue()


def save(fname, obj, cuda=True):
    """``torch.save(fname, obj, 'ascii')``."""
    data = dumps(obj, cuda)
    tmp = "%s.tmp%d" % (fname, os.getpid())
    with open(tmp, "wb") as f:
        f.write(data)
    os.replace(tmp, fname)


# ---------------------------------------------------------------------------------------------- the reference's net files
Net = collections.namedtuple("Net", "arch tower head opt")
Net.__doc__ = """arch 'fast' | 'slow'; tower: [(W (fm, cin, ks, ks), b (fm,)), ...] for FeatureTower; head: [(W (out, in), b (out,)), ...]
for ScorerHead (None for 'fast'); opt: the saved option table (informative: prediction takes its parameters from the command
line, main.lua:56-203)."""


def _weighted(seq):
    if not isinstance(seq, T7Object) or not isinstance(seq.get("modules"), T7Table):
        raise T7Error("t7: expected an nn.Sequential, found %r" % (seq,))
    return [m for m in seq["modules"].array() if isinstance(m, T7Object) and isinstance(m.get("weight"), np.ndarray)
            and m["weight"].size]


def _conv(m):
    w, b = m["weight"], m.get("bias")
    fo, fi = m.get("nOutputPlane"), m.get("nInputPlane")
    kh, kw = m.get("kH"), m.get("kW")
    if None in (fo, fi, kh, kw):
        if w.ndim != 4:
            raise T7Error("t7: %s without nInputPlane / nOutputPlane / kH / kW and a %d-d weight" % (m.typename, w.ndim))
        fo, fi, kh, kw = w.shape
    if w.size != fo * fi * kh * kw or b is None or b.size != fo:
        raise T7Error("t7: %s: weight of %d elements / bias do not match %dx%dx%dx%d" % (m.typename, w.size, fo, fi, kh, kw))
    return (np.ascontiguousarray(w, dtype=np.float32).reshape(fo, fi, kh, kw),
            np.ascontiguousarray(b, dtype=np.float32).reshape(fo))


def _fc(m):
    w, b = m["weight"], m.get("bias")
    if w.ndim != 2 or b is None or b.size != w.shape[0]:
        raise T7Error("t7: %s: expected a (out, in) weight and a bias of `out` elements" % m.typename)
    return np.ascontiguousarray(w, dtype=np.float32), np.ascontiguousarray(b, dtype=np.float32).reshape(-1)


def net_from_object(obj):
    """{net_te, opt} -> arch 'fast' (main.lua:591, 898-900: the trailing StereoJoin module carries no weights and is skipped
    like the reference drops it); {net_te, net_te2, opt} -> arch 'slow' (main.lua:589, 894-896)."""
    if not isinstance(obj, T7Table):
        raise T7Error("t7: a net file holds a table {net_te, [net_te2,] opt}")
    items = obj.array()
    seqs = [v for v in items if isinstance(v, T7Object)]
    opts = [v for v in items if isinstance(v, T7Table)]
    if len(seqs) not in (1, 2) or len(items) != len(seqs) + 1 or len(opts) != 1:
        raise T7Error("t7: expected {net_te, opt} or {net_te, net_te2, opt}, found %d entries" % len(items))
    tower = [_conv(m) for m in _weighted(seqs[0])]
    if not tower:
        raise T7Error("t7: net_te has no convolution layers")
    head = [_fc(m) for m in _weighted(seqs[1])] if len(seqs) == 2 else None
    opt = {k: v for k, v in opts[0].items() if isinstance(k, str)}
    return Net("slow" if head is not None else "fast", tower, head, opt)


def load_net(fname):
    """``torch.load(opt.net_fname, 'ascii')`` + main.lua:893-901 -> :class:`Net`."""
    return net_from_object(load(fname))


def make_net_object(tower, head=None, opt=None):
    """The object ``save_net`` writes (main.lua:587-600) for the given layers: net_te = nn.Sequential of
    cudnn.SpatialConvolution (pad 1) / cudnn.ReLU [+ nn.Normalize2 + nn.StereoJoin for 'fast', main.lua:726-749], net_te2 =
    nn.SpatialConvolution1_fw / cudnn.ReLU ... cudnn.Sigmoid (main.lua:688-695), cleaned as clean_net leaves them (:566-585).
    For exporting weights to the reference and for fixtures."""
    def mod(name, **f):
        t = T7Table(f)
        t["output"] = np.empty((0,), np.float32)                                   # clean_net: empty CudaTensors
        for k in ("finput", "fgradInput", "tmp_in", "tmp_out"):
            t[k] = np.empty((0,), np.float32)
        return T7Object(name, t)

    def seq(mods):
        return mod("nn.Sequential", modules=T7Table(enumerate(mods, 1)))

    fast = head is None
    mods = []
    for i, (w, b) in enumerate(tower):
        w = np.ascontiguousarray(w, dtype=np.float32)
        fo, fi, kh, kw = w.shape
        mods.append(mod("cudnn.SpatialConvolution", nInputPlane=fi, nOutputPlane=fo, kW=kw, kH=kh, dW=1, dH=1,
                        padW=(kw - 1) // 2, padH=(kh - 1) // 2, groups=1, weight=w,
                        bias=np.ascontiguousarray(b, dtype=np.float32).reshape(fo)))
        if not fast or i < len(tower) - 1:
            mods.append(mod("cudnn.ReLU", inplace=True, mode="CUDNN_ACTIVATION_RELU"))
    if fast:
        mods.append(mod("nn.Normalize2"))
        mods.append(mod("nn.StereoJoin", disp_max=1))
    out = [seq(mods)]
    if not fast:
        mods = []
        for i, (w, b) in enumerate(head):
            w = np.ascontiguousarray(w, dtype=np.float32)
            mods.append(mod("nn.SpatialConvolution1_fw", weight=w,
                            bias=np.ascontiguousarray(b, dtype=np.float32).reshape(1, w.shape[0], 1, 1)))
            mods.append(mod("cudnn.ReLU", inplace=True, mode="CUDNN_ACTIVATION_RELU") if i < len(head) - 1
                        else mod("cudnn.Sigmoid", inplace=True, mode="CUDNN_ACTIVATION_SIGMOID"))
        out.append(seq(mods))
    out.append(T7Table(opt or {}))
    return T7Table(enumerate(out, 1))


def save_net(fname, tower, head=None, opt=None):
    save(fname, make_net_object(tower, head, opt))
