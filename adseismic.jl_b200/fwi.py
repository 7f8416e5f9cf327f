"""The parameterisation / inversion layer that sits upstream of the gradient in the reference, on
torch.autograd instead of the TensorFlow graph: the misfit of a shot is a differentiable torch scalar whose backward
is ONE adjoint sweep of libadseis_b200 (the custom-op gradient of the reference, AcousticOneStepCpu.cpp:62-92 and
tf.gradients through src/Core.jl:31-228), so velocity / Lame / density models can be produced by any torch module
(mean-normalised masked variables as in src/IO.jl:172-197, an NN generator as in src/NN.jl, Gaussian-blob sources as
in src/Utils.jl:603-641) and optimised with L-BFGS (src/Optim.jl:135-193).  torch is plumbing here: the numerics of
loss and gradient run in the CUDA library; there is no torch fallback."""
import numpy as np
import torch


def _handoff(t):
    """torch produced `t` on ITS current stream, the library reads it on the context's stream: order the two."""
    if t.is_cuda:
        torch.cuda.current_stream(t.device).synchronize()
    return t if t.is_cuda else t.numpy()


class _AcousticMisfit(torch.autograd.Function):
    """The gradients are computed by the same sweep as the loss, so forward() copies them out of the plan right away:
    a plan that is re-evaluated before backward() runs (several shots or loss terms on one plan, a line-search probe)
    cannot leak the gradient of a later evaluation into this one."""

    @staticmethod
    def forward(ctx, c, srcv, plan):
        cc = c.detach().contiguous().to(torch.float64)
        plan.set_model(_handoff(cc))
        if srcv is not None:
            ss = srcv.detach().contiguous().to(torch.float64)
            plan.set_srcv(_handoff(ss), rows=ss.shape[0])
        plan.gradient()
        ctx.gc = ctx.gs = None
        if ctx.needs_input_grad[0]:
            buf = torch.empty(plan.model_shape, dtype=torch.float64, device=c.device)
            plan.grad_c(out=buf if buf.is_cuda else buf.numpy())
            plan.ctx.sync()
            ctx.gc = buf
        if srcv is not None and ctx.needs_input_grad[1]:
            ctx.gs = torch.from_numpy(plan.grad_srcv()).to(srcv.device)
        ctx.c_like, ctx.s_like = c, srcv
        return torch.tensor(plan.loss(), dtype=torch.float64, device=c.device)

    @staticmethod
    def backward(ctx, gout):
        c, srcv = ctx.c_like, ctx.s_like
        gc = gs = None
        if ctx.needs_input_grad[0] and ctx.gc is not None:
            gc = (gout * ctx.gc).reshape(c.shape).to(c.dtype)
        if srcv is not None and ctx.needs_input_grad[1] and ctx.gs is not None:
            gs = torch.zeros_like(srcv)
            gs[:ctx.gs.shape[0]] = gout * ctx.gs          # rows >= NSTEP of srcv never enter the simulation
        return gc, gs, None


def acoustic_misfit(plan, c, srcv=None):
    """sum((rcvv - obs)^2) of one shot as a differentiable torch scalar.  `plan`: an AcousticPlan whose observed data
    (and srcv, unless given here as a tensor) are already set; `c`: velocity tensor of plan.model_shape (c^2 under the
    MPI convention), on the CPU or on the plan's GPU."""
    return _AcousticMisfit.apply(c, srcv, plan)


class _ElasticMisfit(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rho, lam, mu, srcv, plan):
        arrs = [t.detach().contiguous().to(torch.float64) for t in (rho, lam, mu)]
        plan.set_model(*[_handoff(a) for a in arrs])
        if srcv is not None:
            ss = srcv.detach().contiguous().to(torch.float64)
            plan.set_srcv(_handoff(ss), rows=ss.shape[0])
        mat = any(ctx.needs_input_grad[:3])
        plan.gradient(mat)
        ctx.g = [None, None, None]
        if mat:                                   # copied out now: see _AcousticMisfit
            for k, (t, fn) in enumerate(zip((rho, lam, mu), (plan.grad_rho, plan.grad_lambda, plan.grad_mu))):
                if ctx.needs_input_grad[k]:
                    ctx.g[k] = torch.from_numpy(fn()).to(t.device)
        ctx.gs = torch.from_numpy(plan.grad_srcv()).to(srcv.device) if (srcv is not None and ctx.needs_input_grad[3]) else None
        ctx.like, ctx.s_like = (rho, lam, mu), srcv
        return torch.tensor(plan.loss(), dtype=torch.float64, device=rho.device)

    @staticmethod
    def backward(ctx, gout):
        srcv = ctx.s_like
        out = [None, None, None]
        for k in range(3):
            if ctx.needs_input_grad[k] and ctx.g[k] is not None:
                t = ctx.like[k]
                out[k] = (gout * ctx.g[k]).reshape(t.shape).to(t.dtype)
        gs = None
        if srcv is not None and ctx.needs_input_grad[3] and ctx.gs is not None:
            gs = torch.zeros_like(srcv)
            gs[:ctx.gs.shape[0]] = gout * ctx.gs
        return out[0], out[1], out[2], gs, None


def elastic_misfit(plan, rho, lam, mu, srcv=None):
    """Elastic counterpart: differentiable w.r.t. rho, lambda, mu (model shape of the plan) and srcv.  When no material
    tensor requires a gradient only the source-time-function adjoint runs (no forward history at all)."""
    return _ElasticMisfit.apply(rho, lam, mu, srcv, plan)


def compute_properties(vp, vs, rho):
    """src/Utils.jl:237-243 on tensors: lambda = rho (vp^2 - 2 vs^2), mu = rho vs^2."""
    return rho * (vp * vp - 2.0 * vs * vs), rho * vs * vs, rho


class ConstantOrVariable(torch.nn.Module):
    """constant_or_variable (src/IO.jl:172-197): a trainable field is stored mean-normalised; with a mask only the
    masked cells move (`mask*x_ + x*(1-mask)`), the rest keep the initial values."""

    def __init__(self, x, trainable=False, mask=None):
        super().__init__()
        x = torch.as_tensor(np.asarray(x), dtype=torch.float64)
        self.trainable = bool(trainable)
        if self.trainable:
            self.meanx = float(x.mean())
            self.register_buffer("x0", x / self.meanx)
            self.x_ = torch.nn.Parameter((x / self.meanx).clone())
            self.register_buffer("mask", None if mask is None else torch.as_tensor(np.asarray(mask), dtype=torch.float64))
        else:
            self.register_buffer("x0", x)

    def forward(self):
        if not self.trainable:
            return self.x0
        x_ = self.x_
        if self.mask is not None:
            x_ = self.mask * x_ + self.x0 * (1 - self.mask)
        return x_ * self.meanx


def variable_source(param, x, y, v, sigma=None):
    """variable_source (src/Utils.jl:603-632): a source-time function `v` spread over ALL padded grid cells with a
    Gaussian blob centred at the (differentiable) position (x, y).  Returns (srci, srcj, srcv[len(v), N])."""
    sigma = 1.0 if sigma is None else float(sigma)
    NX2, NY2 = param.NX + 2, param.NY + 2
    ii, jj = np.meshgrid(np.arange(1, NX2 + 1), np.arange(1, NY2 + 1), indexing="ij")
    srci, srcj = ii.reshape(-1).astype(np.int64), jj.reshape(-1).astype(np.int64)
    x, y, v = (torch.as_tensor(t, dtype=torch.float64) for t in (x, y, v))
    xs = torch.as_tensor((srci - 1) * param.DELTAX, dtype=torch.float64)
    ys = torch.as_tensor((srcj - 1) * param.DELTAY, dtype=torch.float64)
    s2 = np.sqrt(2 * sigma)
    magn = 1.0 / (2 * np.pi * sigma) * torch.exp(-(((xs - x) / (s2 * param.DELTAX)) ** 2 + ((ys - y) / (s2 * param.DELTAY)) ** 2))
    return srci, srcj, v.reshape(-1, 1) * magn.reshape(1, -1)


def LBFGS_(loss_fn, params, max_iter=50, callback=None, history_size=10, tolerance_grad=1e-12,
           tolerance_change=1e-14, max_linesearch=20):
    """LBFGS!(sess, loss, grads, vars; callback) (src/Optim.jl:135-193): minimise `loss_fn()` (a closure returning a
    differentiable torch scalar) over `params` with L-BFGS + strong-Wolfe line search; returns the list of losses:
    losses[0] at the starting point, losses[k] after outer iteration k.  callback(params, iteration, loss) after every
    iteration as in the reference.  Every loss/gradient evaluation is one forward+adjoint sweep of the CUDA library."""
    params = list(params)
    opt = torch.optim.LBFGS(params, lr=1.0, max_iter=1, max_eval=max_linesearch + 1, history_size=history_size,
                            line_search_fn="strong_wolfe", tolerance_grad=tolerance_grad,
                            tolerance_change=tolerance_change)
    seen = {}          # parameter fingerprint -> loss, for the points the line search has evaluated

    def key():
        return hash(tuple(p.detach().cpu().numpy().tobytes() for p in params))

    def closure():
        opt.zero_grad()
        loss = loss_fn()
        loss.backward()
        seen[key()] = float(loss.detach())
        return loss

    def current():
        k = key()
        if k not in seen:
            with torch.enable_grad():
                closure()
        return seen[k]

    losses = []
    for it in range(max_iter):
        seen.clear()
        l0 = float(opt.step(closure).detach())   # returns the loss at the START of the step
        if not losses:
            losses.append(l0)
        losses.append(current())             # the accepted point was evaluated by the line search
        if callback is not None:
            callback(params, it, losses[-1])
        if abs(losses[-2] - losses[-1]) <= tolerance_change * max(1.0, abs(losses[-1])):
            break
    return losses
