"""Python face of the library, mirroring the reference's operator interface for the
hot path (same names, argument meaning and error behaviour):

    exec = B200Executor.create(0)                 # gko::CudaExecutor::create
    A = Csr.create(exec, (n, m), values, col_idxs, row_ptrs)   # gko::matrix::Csr::create
    A.apply(b, x)                                 # LinOp::apply(b, x):          x = A b
    A.apply(alpha, b, beta, x)                    # LinOp::apply(alpha,b,beta,x) x = aAb + bx

Device memory and streams come from torch (plumbing); every computation goes through the
C ABI of include/ginkgo_b200.h into hand-written sm_100a kernels.  Operands that live on the
host are cloned to the executor for the call and the result copied back, exactly like
LinOp::apply's make_temporary_clone (include/ginkgo/core/base/lin_op.hpp:129-215).
There is no CPU fallback."""
import ctypes

import torch

from . import _lib

_VT = {torch.float64: "f64", torch.float32: "f32"}
_IT = {torch.int32: "i32", torch.int64: "i64"}


class DimensionMismatch(ValueError):
    """gko::DimensionMismatch"""


class NotSupported(TypeError):
    """gko::NotSupported"""


class B200Executor:
    """gko::CudaExecutor analogue: one device + one stream + the C-ABI context."""

    def __init__(self, device_id, stream):
        self.device_id = device_id
        self.device = torch.device("cuda", device_id)
        self.stream = stream
        self._l = _lib.lib()
        ctx = ctypes.c_void_p()
        _lib.check(self._l.b200_ctx_create(device_id, stream.cuda_stream, ctypes.byref(ctx)))
        self.ctx = ctx

    @staticmethod
    def create(device_id=0, stream=None):
        if not torch.cuda.is_available():
            raise _lib.B200Error("B200Executor needs a CUDA device (no CPU fallback)")
        with torch.cuda.device(device_id):
            stream = stream or torch.cuda.Stream(device_id)
        return B200Executor(device_id, stream)

    def synchronize(self):
        _lib.check(self._l.b200_synchronize(self.ctx))

    def launch_count(self):
        return self._l.b200_ctx_launch_count(self.ctx)

    def num_sms(self):
        return self._l.b200_ctx_num_sms(self.ctx)

    def run(self, name, *args):
        """exec->run(make_<op>(...)): enqueue one C-ABI call on the executor's stream"""
        _lib.call(name, self.ctx, *args)

    def __del__(self):
        try:
            self._l.b200_ctx_destroy(self.ctx)
        except Exception:
            pass


class Dense:
    """gko::matrix::Dense: row-major values with a row stride."""

    def __init__(self, exec_, values, size=None, stride=None):
        if values.dim() == 1:
            values = values.reshape(-1, 1)
        self.exec = exec_
        self.values = values
        self.size = tuple(size) if size is not None else tuple(values.shape)
        self.stride = stride if stride is not None else values.stride(0) if values.shape[0] > 1 \
            else values.shape[1]
        if values.dtype not in _VT:
            raise NotSupported("value type %s" % values.dtype)

    @staticmethod
    def create(exec_, size, values=None, stride=None, dtype=torch.float64):
        rows, cols = size
        stride = cols if stride is None else stride
        if values is None:
            with torch.cuda.stream(exec_.stream):
                values = torch.empty((rows, stride), dtype=dtype, device=exec_.device)
        return Dense(exec_, values, (rows, cols), stride)

    @property
    def vt(self):
        return _VT[self.values.dtype]

    @property
    def on_device(self):
        return self.values.is_cuda

    def to_device(self, exec_):
        if self.on_device:
            return self
        with torch.cuda.stream(exec_.stream):
            v = self.values.to(exec_.device, non_blocking=True)
        return Dense(exec_, v, self.size, self.stride)


class _SparseBase:
    def _check(self, b, x):
        if b.size[0] != self.size[1] or x.size[0] != self.size[0] or b.size[1] != x.size[1]:
            raise DimensionMismatch("apply: A %s, b %s, x %s" % (self.size, b.size, x.size))
        if b.vt != self.vt or x.vt != self.vt:
            raise NotSupported("mixed precision apply is not compiled (GINKGO_MIXED_PRECISION off)")

    def apply(self, *args):
        """apply(b, x) or apply(alpha, b, beta, x)"""
        if len(args) == 2:
            alpha = beta = None
            b, x = args
        elif len(args) == 4:
            alpha, b, beta, x = args
        else:
            raise TypeError("apply takes (b, x) or (alpha, b, beta, x)")
        self._check(b, x)
        e = self.exec
        with torch.cuda.stream(e.stream):
            bd = b.to_device(e)
            xd = x.to_device(e) if (alpha is not None or x.on_device) else Dense.create(
                e, x.size, dtype=x.values.dtype)
            ad = alpha.to_device(e) if alpha is not None else None
            btd = beta.to_device(e) if beta is not None else None
            self._apply_impl(ad, bd, btd, xd)
            if not x.on_device:
                x.values.copy_(xd.values, non_blocking=True)
                e.stream.synchronize()
        return x


class Csr(_SparseBase):
    """gko::matrix::Csr<ValueType, IndexType> on a B200Executor."""

    def __init__(self, exec_, size, values, col_idxs, row_ptrs):
        self.exec = exec_
        self.size = tuple(size)
        self.values, self.col_idxs, self.row_ptrs = values, col_idxs, row_ptrs
        self.nnz = values.numel()
        if row_ptrs.numel() != size[0] + 1:
            raise DimensionMismatch("row_ptrs must have num_rows + 1 entries")
        if values.dtype not in _VT or col_idxs.dtype not in _IT or row_ptrs.dtype != col_idxs.dtype:
            raise NotSupported("value/index type")
        self.vt, self.it = _VT[values.dtype], _IT[col_idxs.dtype]
        self._plan = None

    @staticmethod
    def create(exec_, size, values, col_idxs, row_ptrs):
        with torch.cuda.stream(exec_.stream):
            v, c, r = (t.to(exec_.device) for t in (values, col_idxs, row_ptrs))
        return Csr(exec_, size, v.contiguous(), c.contiguous(), r.contiguous())

    def plan(self):
        # the `srow` analogue: computed once per matrix (strategy->process in the reference)
        if self._plan is None:
            p = ctypes.c_void_p()
            fn = getattr(self.exec._l, "b200_csr_plan_create_%s_%s" % (self.vt, self.it))
            _lib.check(fn(self.exec.ctx, self.size[0], self.nnz, self.row_ptrs.data_ptr(),
                          ctypes.byref(p)))
            self._plan = p
        return self._plan

    def _apply_impl(self, alpha, b, beta, x):
        sfx = "_%s_%s" % (self.vt, self.it)
        common = (self.plan(), self.size[0], self.size[1], self.nnz, self.row_ptrs, self.col_idxs,
                  self.values)
        if alpha is None:
            self.exec.run("b200_csr_spmv" + sfx, *common, b.values, b.stride, b.size[1], x.values,
                          x.stride)
        else:
            self.exec.run("b200_csr_advanced_spmv" + sfx, *common, alpha.values, b.values, b.stride,
                          b.size[1], beta.values, x.values, x.stride)

    def __del__(self):
        try:
            if self._plan is not None:
                self.exec._l.b200_csr_plan_destroy(self._plan)
        except Exception:
            pass
