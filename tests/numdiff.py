"""Numeric-differentiation harness (oracle style (c) of SURVEY §4;
/root/reference/tests/unit/gd_numdiff.py:43-156): differentiate the forward unit
numerically in float64 and compare against the GD unit's analytic outputs."""
import numpy

from veles.znicz_b200.core.config import root
from veles.znicz_b200.core.memory import Array
from veles.znicz_b200.core.workflow import DummyWorkflow


def numeric_grad(f, x, h=1e-6):
    """d f() / d x for scalar-valued f, central 5-point stencil, in place."""
    g = numpy.zeros_like(x)
    flat = x.reshape(-1)
    gf = g.reshape(-1)
    for i in range(flat.size):
        old = flat[i]
        vals = []
        for d in (2 * h, h, -h, -2 * h):
            flat[i] = old + d
            vals.append(f())
        flat[i] = old
        gf[i] = (-vals[0] + 8 * vals[1] - 8 * vals[2] + vals[3]) / (12 * h)
    return g


def check_pair(fwd_cls, gd_cls, inp, fwd_kwargs=None, gd_kwargs=None, tol=1e-5,
               link=("weights", "bias"), extra_links=(), seed=3):
    """Builds fwd+gd on the numpy backend in float64 and checks err_input,
    gradient_weights, gradient_bias against numeric derivatives of
    L = sum(output * R) for a fixed random R."""
    root.common.engine.precision_type = "double"
    try:
        rs = numpy.random.RandomState(seed)
        wf = DummyWorkflow()
        fwd = fwd_cls(wf, **(fwd_kwargs or {}))
        fwd.input = Array(inp.astype(numpy.float64).copy())
        fwd.initialize(device=None)
        fwd.run()
        r = rs.uniform(-1, 1, fwd.output.shape)

        def loss():
            fwd.run()
            return float((fwd.output.mem * r).sum())

        kw = dict(learning_rate=1.0, learning_rate_bias=1.0, weights_decay=0.0,
                  weights_decay_bias=0.0, gradient_moment=0.0, apply_gradient=False)
        kw.update(gd_kwargs or {})
        gd = gd_cls(wf, **kw)
        gd.err_output = Array(r.copy())
        gd.input = fwd.input
        gd.output = fwd.output
        for a in link:
            if getattr(fwd, a, None) is not None:
                setattr(gd, a, getattr(fwd, a))
        for a in extra_links:
            setattr(gd, a, getattr(fwd, a))
        gd.initialize(device=None)
        fwd.run()
        gd.err_output.mem[...] = r   # GD units may scale err_output in place
        gd.run()
        res = {}
        ng = numeric_grad(loss, fwd.input.mem)
        res["err_input"] = float(numpy.abs(ng - gd.err_input.mem.reshape(ng.shape)).max())
        assert res["err_input"] < tol, ("err_input", res)
        if "weights" in link and getattr(fwd, "weights", None):
            ngw = numeric_grad(loss, fwd.weights.mem)
            res["gw"] = float(numpy.abs(
                ngw - gd.gradient_weights.mem.reshape(ngw.shape)).max())
            assert res["gw"] < tol, ("gradient_weights", res)
            if getattr(fwd, "bias", None) and fwd.include_bias:
                ngb = numeric_grad(loss, fwd.bias.mem)
                res["gb"] = float(numpy.abs(ngb - gd.gradient_bias.mem).max())
                assert res["gb"] < tol, ("gradient_bias", res)
        return res
    finally:
        root.common.engine.precision_type = "float"
