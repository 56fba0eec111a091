"""dgl.function builtins used by the reference (gvp.py:474,490,496): u_sub_v, copy_e, mean, sum."""


class _BinaryMsg:
    def __init__(self, lhs, rhs, out, op):
        self.lhs, self.rhs, self.out, self.op = lhs, rhs, out, op


class _CopyE:
    def __init__(self, field, out):
        self.field, self.out = field, out


class _Reduce:
    def __init__(self, kind, msg, out):
        self.kind, self.msg, self.out = kind, msg, out


def u_sub_v(lhs, rhs, out):
    return _BinaryMsg(lhs, rhs, out, lambda a, b: a - b)


def copy_e(e, out):
    return _CopyE(e, out)


def mean(msg, out):
    return _Reduce("mean", msg, out)


def sum(msg, out):  # noqa: A001 - mirrors dgl.function.sum
    return _Reduce("sum", msg, out)
