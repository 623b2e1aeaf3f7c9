"""Two interchangeable backends with the Jets API surface, so every parity scenario is written
once and run through (a) the numpy oracle and (b) the CUDA library via its C ABI."""
import numpy as np


class OracleBackend:
    name = "oracle"

    def __init__(self):
        from oracle import jets_oracle as J
        self.J = J
        for k in ("JetSpace", "JetBSpace", "JopDiagonal", "JopScale", "JopPointwise", "JopStencil",
                  "JopDense", "JopZeroBlock", "blockop", "compose", "adjoint", "jacobian", "jacobian_",
                  "mul_", "dot", "norm", "extrema", "dot_product_test", "linearity_test", "to_matrix",
                  "getblock", "nblocks", "domain", "range_", "zeros", "ones", "iszero", "isblockop",
                  "size", "shape", "setblock_", "getblock_", "fill_", "JopRestriction"):
            setattr(self, k, getattr(J, k))

    def arr(self, x, R):
        """numpy -> backend array living in space R (a private copy)."""
        x = np.array(x, dtype=R.T, copy=True)
        if isinstance(R, self.J.JetBSpace):
            return self.J.reshape(x.reshape(-1, order="F"), R)
        return np.asfortranarray(x.reshape(R.n, order="F"))

    def host(self, x):
        if isinstance(x, self.J.BlockArray):
            return self.J.to_array(x)
        return np.asarray(x)

    def block_ranges(self, x):
        return list(x.indices)

    def lincomb(self, terms):
        xs = [x for _, x in terms]
        cs = [xs[0].dtype.type(c) for c, _ in terms]

        def f(*bl):
            out = cs[0] * bl[0]
            for c, b in zip(cs[1:], bl[1:]):
                out = out + c * b
            return out
        if isinstance(xs[0], self.J.BlockArray):
            return self.J.bmap(f, *xs)
        return f(*xs)

    def hadamard(self, x, y):
        return x * y


class DeviceBackend:
    name = "device"

    def __init__(self):
        import jets_b200 as B
        self.B = B
        for k in ("JetSpace", "JetBSpace", "JopDiagonal", "JopScale", "JopPointwise", "JopStencil",
                  "JopDense", "JopZeroBlock", "blockop", "compose", "adjoint", "jacobian", "jacobian_",
                  "mul_", "dot", "norm", "extrema", "dot_product_test", "linearity_test", "to_matrix",
                  "getblock", "nblocks", "domain", "range_", "zeros", "ones", "iszero", "isblockop",
                  "size", "shape", "setblock_", "getblock_", "fill_", "JopRestriction"):
            setattr(self, k, getattr(B, k))

    def arr(self, x, R):
        return self.B.to_device(np.asarray(x, dtype=R.T), R)

    def host(self, x):
        return self.B.to_array(x)

    def block_ranges(self, x):
        return [x.block_range(i + 1) for i in range(self.B.nblocks(x))]

    def lincomb(self, terms):
        out = self.B.similar(terms[0][1])
        return self.B.lincomb_(out, terms)

    def hadamard(self, x, y):
        return x * y
