"""Host-side mirror of the reference's Wolfram Language API for the nested-sampling path.

Same function names, option names, result keys and error behaviour as the package:
  defineInferenceProblem   BS:148-308      generateStartingPoints  BS:1042-1097
  nestedSampling           BS:1099-1136    parallelNestedSampling  BS:1317-1371
  evidenceSampling         BS:1158-1291    combineRuns             BS:1293-1315
  inferenceObject          BU:107-138      defineGaussianProcess   GP:201-330
All numerics run in libbinest.so (CUDA); this file only marshals arguments and assembles associations.
The Wolfram Language host package (wl/BayesianInferenceB200.wl) is the same layer in the reference's own
language; it cannot be executed in this image (no Wolfram Engine), so tests drive this mirror.

A symbolic "GeneratingDistribution" is replaced by a descriptor from the fixed operator table
(SURVEY.md Appendix A mapping); anything else fails like the reference does when LogLikelihood does not
evaluate: message + inferenceObject[$Failed] (BS:456-459, 308).  There is no CPU fallback.
"""
from __future__ import annotations

import warnings
from dataclasses import dataclass

import numpy as np

from . import configs as _cfg

FAILED = "$Failed"
LOGZERO_HOST = -1.7976931348623157e308  # $MachineLogZero, BU:47


# ----------------------------------------------------------------------------------------------------
# inferenceObject (BU:107-138): an association wrapper with obj["Key"] access
class inferenceObject:
    def __init__(self, assoc):
        self._assoc = assoc

    def __getitem__(self, key):
        if self.failed:
            raise KeyError(key)
        v = self._assoc.get(key, None)
        if v is None and key not in self._assoc:
            return Missing("KeyAbsent", key)
        return v

    def get(self, key, default=None):
        return default if self.failed else self._assoc.get(key, default)

    def keys(self):
        return [] if self.failed else [k for k in self._assoc if not k.startswith("_")]

    def Normal(self):  # BU:125
        return dict(self._assoc) if not self.failed else FAILED

    @property
    def failed(self):
        return not isinstance(self._assoc, dict)

    def __contains__(self, key):
        return (not self.failed) and key in self._assoc

    def __repr__(self):
        if self.failed:
            return "inferenceObject[$Failed]"
        return f"inferenceObject[<|{', '.join(self.keys())}|>]"


@dataclass(frozen=True)
class Missing:
    reason: str
    key: str = ""

    def __bool__(self):
        return False


def inferenceObjectQ(x):
    return isinstance(x, inferenceObject) and not x.failed


# ----------------------------------------------------------------------------------------------------
# distribution descriptors: the fixed operator table
@dataclass(frozen=True)
class Polynomial:
    """Sum_j coefficients[j] * variable^j"""
    variable: str
    coefficients: tuple


@dataclass(frozen=True)
class NormalDistribution:
    mean: object  # parameter name, number (priors) or Polynomial
    sd: object


@dataclass(frozen=True)
class UniformDistribution:
    lo: float
    hi: float


@dataclass(frozen=True)
class CategoricalSoftmax:
    """Softmax classification with reference class K: logits z_k = b_k + w_k . x (k < K), z_K = 0.
    params: K-1 blocks (w_1..w_F, b) of parameter names."""
    params: tuple
    n_classes: int = 3


@dataclass(frozen=True)
class GeometricBrownianMotionProcess:
    mu: str
    sigma: str
    x0: float = None


@dataclass(frozen=True)
class SquaredExponentialGP:
    sigma_f: str
    ell: str
    sigma_n: str


class DefinitionError(ValueError):
    pass


_PARAM_MSG = "defineInferenceProblem::parameters"


def _param_normal_form(params):
    """paramNormalForm BS:133-145: {sym, lo, hi} triples."""
    out = []
    for p in params:
        if isinstance(p, str):
            out.append((p, -np.inf, np.inf))
        else:
            name, lo, hi = p
            out.append((str(name), float(lo), float(hi)))
    return out


def _prior_spec(prior, params):
    """ignorancePrior BS:25-64: list entries "LocationParameter" | "ScaleParameter" | distribution."""
    kinds, p0, p1 = [], [], []
    if not isinstance(prior, (list, tuple)) or len(prior) != len(params):
        raise DefinitionError("PriorDistribution must be a list with one entry per parameter")
    for spec, (_, lo, hi) in zip(prior, params):
        if spec == "LocationParameter" or isinstance(spec, UniformDistribution):
            kinds.append(_cfg.PRIOR_UNIFORM); p0.append(0.0); p1.append(1.0)
        elif spec == "ScaleParameter":
            kinds.append(_cfg.PRIOR_SCALE); p0.append(0.0); p1.append(1.0)
        elif isinstance(spec, NormalDistribution) and np.isscalar(spec.mean) and np.isscalar(spec.sd):
            kinds.append(_cfg.PRIOR_NORMAL_TRUNC); p0.append(float(spec.mean)); p1.append(float(spec.sd))
        else:
            raise DefinitionError(f"defineInferenceProblem::prior: unsupported prior {spec!r}")
    return kinds, p0, p1


def _operator_from_distribution(dist, data, names, indep):
    """Map a "GeneratingDistribution" pattern to (op id, iparam, inputs, outputs) — SURVEY Appendix A."""
    if isinstance(dist, NormalDistribution) and isinstance(dist.mean, str) and isinstance(dist.sd, str):
        x = np.asarray(data, dtype=np.float64)
        if isinstance(data, tuple) or x.ndim > 2 or (x.ndim == 2 and x.shape[1] != 1):
            raise DefinitionError("NormalDistribution[mu, sigma] needs vector data")
        if names != [dist.mean, dist.sd]:
            raise DefinitionError("Parameters must be ordered {mu, sigma}")
        return _cfg.OP_GAUSSIAN_IID, (0, 0, 0, 0), x.reshape(-1, 1), None
    if isinstance(dist, NormalDistribution) and isinstance(dist.mean, Polynomial):
        if not (isinstance(data, tuple) and len(data) == 2):
            raise DefinitionError("regression data must be (inputs, outputs)")
        poly = dist.mean
        if indep is None or list(indep) != [poly.variable]:
            raise DefinitionError("IndependentVariables must name the polynomial's variable")
        deg = len(poly.coefficients) - 1
        if names != list(poly.coefficients) + [dist.sd]:
            raise DefinitionError("Parameters must be ordered {c_0..c_deg, sigma}")
        return _cfg.OP_POLYREG, (deg, 0, 0, 0), np.asarray(data[0], float).reshape(-1, 1), np.asarray(data[1], float).reshape(-1, 1)
    if isinstance(dist, CategoricalSoftmax):
        x, y = data
        x = np.asarray(x, float)
        x = x.reshape(-1, 1) if x.ndim == 1 else x
        if names != list(dist.params):
            raise DefinitionError("Parameters must be ordered as the softmax blocks (w_1..w_F, b)")
        return _cfg.OP_LOGISTIC, (0, dist.n_classes, 0, 0), x, np.asarray(y, float).reshape(-1, 1)
    if isinstance(dist, GeometricBrownianMotionProcess):
        t, v = data  # TemporalData adaptor BS:511-515
        if names != [dist.mu, dist.sigma]:
            raise DefinitionError("Parameters must be ordered {mu, sigma}")
        return _cfg.OP_GBM, (0, 0, 0, 0), np.asarray(t, float).reshape(-1, 1), np.asarray(v, float).reshape(-1, 1)
    if isinstance(dist, SquaredExponentialGP):
        x, y = data
        x = np.asarray(x, float)
        x = x.reshape(-1, 1) if x.ndim == 1 else x
        return _cfg.OP_GP_SE, (0, 0, x.shape[1], 0), x, np.asarray(y, float).reshape(-1, 1)
    # the reference would try LogLikelihood[dist, ...] symbolically (BS:452-459); only the table runs on the GPU
    raise DefinitionError(f"defineInferenceProblem::logLike: {dist!r} is not in the GPU operator table")


def _backend(backend):
    if backend is not None:
        return backend
    from . import engine
    return engine


def defineInferenceProblem(rules=None, _backend_override=None, **kw):
    """BS:148-308.  Keys: "Data", "Parameters", "PriorDistribution", "GeneratingDistribution",
    "IndependentVariables".  Returns inferenceObject[...] or inferenceObject[$Failed] (+ a warning carrying the
    reference's message tag).

    Extra key "DataSharding" (no reference counterpart; SURVEY §8e "very large N"): None (default), a backend Comm,
    or "Automatic" = one shard per torch.distributed rank.  Every rank passes the full "Data"; only its row block
    is uploaded, and "LogLikelihoodFunction" / nestedSampling become collectives that all ranks must call alike.
    The GP operator is batch-sharded instead (§8e row 2): the data stay replicated and every batch of parameter
    vectors is split across the ranks."""
    a = dict(rules or {})
    a.update(kw)
    try:
        for k in ("Data", "Parameters", "GeneratingDistribution", "PriorDistribution"):
            if k not in a:
                raise DefinitionError(f"defineInferenceProblem::insuffInfo: missing {k}")  # BS:148-152
        params = _param_normal_form(a["Parameters"])
        names = [p[0] for p in params]
        kinds, p0, p1 = _prior_spec(a["PriorDistribution"], params)
        op, iparam, inputs, outputs = _operator_from_distribution(a["GeneratingDistribution"], a["Data"], names,
                                                                   a.get("IndependentVariables"))
        be = _backend(_backend_override)
        comm = a.get("DataSharding")
        if isinstance(comm, str):
            if comm != "Automatic":
                raise DefinitionError(f"defineInferenceProblem::insuffInfo: bad DataSharding value {comm!r}")
            import torch.distributed as dist
            comm = None
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                comm = be.Comm(dist.get_rank(), dist.get_world_size())
        extra = {} if comm is None else {"comm": comm}
        if comm is not None and op == _cfg.OP_GP_SE:
            extra["shard"] = "batch"  # a covariance matrix does not shard by rows: the theta batch is split instead
        prob = be.Problem(op, inputs, outputs, iparam, kinds, [p[1] for p in params], [p[2] for p in params], p0, p1,
                          **extra)
        a["DataSharding"] = comm
        # BS:276-298: both functions must return real machine numbers on random points of the box
        test = prob.sample_prior(100, seed=20260, run_id=0)
        ll, lp = prob.loglike(test), prob.logprior(test)
        if not (np.all(np.isfinite(ll)) and np.all(np.isfinite(lp))):
            raise DefinitionError("defineInferenceProblem::failed: functions are not numeric on the parameter box")
    except DefinitionError as e:
        warnings.warn(str(e))
        return inferenceObject(FAILED)
    a.update({
        "Parameters": params, "ParameterSymbols": names,
        "LogLikelihoodFunction": prob.loglike,   # Listable, BS:499
        "LogPriorPDFFunction": prob.logprior,    # BS:410-426
        "_problem": prob, "_backend": be,
    })
    return inferenceObject(a)


def defineGaussianProcess(data, kernel: SquaredExponentialGP, parameters, prior, **rest):
    """GP:201-330 for the squared-exponential kernel + nugget (the operator of BASELINE config C5)."""
    obj = defineInferenceProblem(Data=data, GeneratingDistribution=kernel, Parameters=parameters,
                                 PriorDistribution=prior, **rest)
    if not inferenceObjectQ(obj):
        return obj
    a = obj.Normal()
    x = np.asarray(data[0], float)
    a["Data"] = (x.reshape(-1, 1) if x.ndim == 1 else x, np.asarray(data[1], float).reshape(-1, 1))  # dataNormalForm
    # GP:312-320: the model functions live in the device operator; the key marks the object as a GP problem
    a["GaussianProcessData"] = {"ModelFunctions": {"KernelFunction": kernel, "NuggetFunction": kernel.sigma_n,
                                                   "MeanFunction": 0.0}}
    return inferenceObject(a)


class MixtureDistribution:
    """MixtureDistribution[weights, {NormalDistribution[mu_m, sigma_m]..}] (GP:357): the posterior predictive at
    one input, a weighted mixture over the samples of the run."""

    def __init__(self, weights, means, sds):
        w = np.asarray(weights, float)
        self.weights, self.means, self.sds = w / w.sum(), np.asarray(means, float), np.asarray(sds, float)

    def mean(self):
        return float(self.weights @ self.means)

    def variance(self):
        m = self.mean()
        return float(self.weights @ (self.sds**2 + (self.means - m) ** 2))

    def sd(self):
        return float(np.sqrt(self.variance()))

    def pdf(self, x):
        x = np.asarray(x, float)[..., None]
        z = (x - self.means) / self.sds
        return (self.weights * np.exp(-0.5 * z * z) / (self.sds * np.sqrt(2 * np.pi))).sum(-1)

    def cdf(self, x):
        from math import erf
        x = np.asarray(x, float)[..., None]
        z = (x - self.means) / (self.sds * np.sqrt(2.0))
        return (self.weights * 0.5 * (1.0 + np.vectorize(erf)(z))).sum(-1)

    def quantile(self, q):
        lo, hi = float((self.means - 10 * self.sds).min()), float((self.means + 10 * self.sds).max())
        for _ in range(200):
            mid = 0.5 * (lo + hi)
            lo, hi = (mid, hi) if self.cdf(mid) < q else (lo, mid)
        return 0.5 * (lo + hi)


class CategoricalMixture:
    """MixtureDistribution[weights, {CategoricalDistribution[p_m]..}]: class probabilities are the weighted mean."""

    def __init__(self, weights, probs):
        w = np.asarray(weights, float)
        self.weights, self.probs = w / w.sum(), np.asarray(probs, float)

    def probabilities(self):
        return self.weights @ self.probs

    def mode(self):
        return int(np.argmax(self.probabilities()))


class ParameterMixture:
    """MixtureDistribution[weights, dist /@ points] for a generating distribution without independent variables
    (BS:1421-1435): the components are the distribution at every sample's parameters."""

    def __init__(self, weights, points, distribution, names):
        w = np.asarray(weights, float)
        self.weights, self.points, self.distribution, self.names = w / w.sum(), np.asarray(points, float), distribution, names


class Predictive(dict):
    """Association key -> mixture (BS:1464-1483) plus the arrays behind it: .inputs (Q, F), .weights (M,),
    .components (M, Q, C)."""


def predictiveDistribution(obj, inputs=None, keys=None, point_estimate=None, _backend_override=None):
    """BS:1373-1483.  point_estimate: None (all samples, weighted by "CrudePosteriorWeight"), "MaximumLikelihood"
    (BS:1388-1402) or "MAP" (BS:1404-1418).  Without inputs: the mixture over the samples' parameters (i.i.d. data,
    BS:1420-1435); with inputs (and optional keys, BS:1437-1448): one mixture per input of a regression problem."""
    if not inferenceObjectQ(obj):
        return FAILED
    a = obj.Normal()
    if "Samples" not in a:
        warnings.warn("predictiveDistribution::unsampled: Posterior has not been sampled yet")  # BS:1375-1380
        return FAILED
    if "GeneratingDistribution" not in a:
        warnings.warn("predictiveDistribution::MissGenDist: No generating distribution specified")  # BS:1381-1387
        return FAILED
    S = a["Samples"]
    pts = np.asarray(S["Point"], float)
    if point_estimate in ("MaximumLikelihood", "MAP"):
        score = S["LogLikelihood"] + (S["LogPriorPDF"] if point_estimate == "MAP" else 0.0)
        i = int(np.argmax(score))
        pts, w = pts[i:i + 1], np.ones(1)
    elif point_estimate is None:
        if "CrudePosteriorWeight" not in S:
            return FAILED
        w = np.asarray(S["CrudePosteriorWeight"], float)
    else:
        raise TypeError(f"unknown point estimate {point_estimate!r}")
    dist, names = a["GeneratingDistribution"], a["ParameterSymbols"]
    regression = isinstance(a["Data"], tuple) and a.get("IndependentVariables") is not None or isinstance(dist, CategoricalSoftmax)
    if inputs is None:
        if regression:
            return FAILED  # BS:1420: this form needs ListQ data (no independent variables)
        if isinstance(dist, NormalDistribution):
            return MixtureDistribution(w, pts[:, 0], pts[:, 1])
        return ParameterMixture(w, pts, dist, names)
    if not regression:
        return FAILED
    x = np.asarray(inputs, float)
    x = x.reshape(-1, 1) if x.ndim == 1 else x  # dataNormalForm BS:1437-1441
    if x.ndim != 2 or not np.all(np.isfinite(x)):
        return FAILED
    if keys is None:
        keys = [float(r[0]) if x.shape[1] == 1 else tuple(float(v) for v in r) for r in x]  # BS:1443-1446
    if len(keys) != x.shape[0]:
        return FAILED  # BS:1459
    comp = a["_problem"].predictive_components(pts, x)
    out = Predictive()
    for q, k in enumerate(keys):
        if isinstance(dist, CategoricalSoftmax):
            out[k] = CategoricalMixture(w, comp[:, q, :])
        else:
            out[k] = MixtureDistribution(w, comp[:, q, 0], comp[:, q, 1])
    out.inputs, out.weights, out.components = x, w / w.sum(), comp
    return out


# ----------------------------------------------------------------------------------------------------
# posterior sampler
class MarkovChain:
    """The chain object of createMCMCChain: chain["AcceptanceRate"], chain["StateData"] = (x, t, mean, cov)."""

    def __init__(self, backend_chain, names):
        self._c, self.names = backend_chain, names

    def __getitem__(self, key):
        st = self._c.state()
        one = self._c.n_chains == 1
        if key == "AcceptanceRate":
            r = st["accepted"] / np.maximum(st["t"] - 1, 1)
            return float(r[0]) if one else r
        if key == "StateData":
            return tuple(v[0] if one else v for v in (st["x"], st["t"], st["mean"], st["cov"]))
        return Missing("KeyAbsent")


def createMCMCChain(obj, startPt=None, _backend_override=None, **opts):
    """BS:651-701.  Options "InitialCovariance" (number, vector or matrix; default 1, BS:676-683) and
    "CovarianceLearnDelay" (default 20), plus "Seed" and "Chains" (independent chains advanced in lock-step on the
    GPU; 1 is the reference).  startPt defaults to the first of obj["StartingPoints"] (BS:657-658); without either
    the reference issues createMCMCChain::start and returns inferenceObject[$Failed]."""
    o = {"InitialCovariance": 1, "CovarianceLearnDelay": 20, "Seed": 1, "Chains": 1}
    bad = [k for k in opts if k not in o]
    if bad:
        raise TypeError(f"Unknown option(s) {bad}")
    o.update(opts)
    if not inferenceObjectQ(obj):
        return inferenceObject(FAILED)
    a = obj.Normal()
    d = len(a["Parameters"])
    nch = int(o["Chains"])
    if startPt is None:
        sp = a.get("StartingPoints")
        if sp is None or np.ndim(sp) != 2:
            warnings.warn("createMCMCChain::start: Please specify a starting point")  # BS:651-655
            return inferenceObject(FAILED)
        startPt = np.asarray(sp, float)[:nch]
    start = np.atleast_2d(np.asarray(startPt, float))
    if start.shape != (nch, d):
        return inferenceObject(FAILED)
    ic = o["InitialCovariance"]
    if np.ndim(ic) == 0 and isinstance(ic, (int, float, np.integer, np.floating)):
        cov = np.eye(d) * float(ic)                       # n -> DiagonalMatrix[ConstantArray[n, dim]]
    elif np.ndim(ic) == 1 and len(ic) == d:
        cov = np.diag(np.asarray(ic, float))              # vector -> DiagonalMatrix
    elif np.ndim(ic) == 2 and np.shape(ic) == (d, d):
        cov = np.asarray(ic, float)
    else:
        cov = np.eye(d)                                   # anything else -> identity (BS:681)
    delay = o["CovarianceLearnDelay"]
    delay = int(delay) if isinstance(delay, (int, np.integer)) else 20  # BS:684-690
    be = _backend(_backend_override or a.get("_backend"))
    return MarkovChain(be.Chain(a["_problem"], start, cov, delay, int(o["Seed"])), a["ParameterSymbols"])


def iterateMCMC(chain, spec):
    """Statistics`MCMC`MarkovChainIterate (BS:703): spec = n -> the next n states; {n, thin} -> n states, one every
    `thin` steps (the forms used at BS:729 and BS:1089-1090).  One chain: (n, d); several: (n, chains, d)."""
    if isinstance(spec, (int, np.integer)):
        n, thin = int(spec), 1
    else:
        n, thin = int(spec[0]), int(spec[1])
    out = chain._c.iterate(n * thin)[thin - 1::thin]
    return out[:, 0, :] if chain._c.n_chains == 1 else out


# ----------------------------------------------------------------------------------------------------
# Laplace evidence: mode + Hessian of the log posterior on the GPU operators
def laplaceLogEvidence(maximum, precisionMatrix):
    """LA:22-30: max + (n Log[2 Pi] - Log[Det[precision]]) / 2, Missing[] unless the determinant is positive."""
    P = np.atleast_2d(np.asarray(precisionMatrix, float))
    sign, logdet = np.linalg.slogdet(P)
    if not (sign > 0 and np.isfinite(logdet)):
        return Missing("NotPositiveDefinite")
    return float(maximum + 0.5 * (P.shape[0] * np.log(2 * np.pi) - logdet))


def approximateEvidence(obj, InitialGuess=None, MaxIterations=50, _backend_override=None):
    """approximateEvidence of the reference (LA:177-238) on an inferenceObject: maximise the unnormalised log
    posterior "LogLikelihoodFunction" + "LogPriorPDFFunction" inside the parameter box, take minus its Hessian at the
    maximum as the precision matrix, and return <|"LogEvidence", "Maximum", "Mean", "PrecisionMatrix", "Parameters"|>
    (LA:219-234).  The reference differentiates the symbolic density (FindMaximum + numericD "Hessian"); the operators
    here are device code, so gradient and Hessian are central differences — all 2 d^2 + 2 d + 1 stencil points of a Newton
    iteration go through ONE batched "LogLikelihoodFunction" call, the shape the GPU operators are built for.
    InitialGuess: a point, or None = the best sample of a finished run / of 4096 prior draws."""
    if not inferenceObjectQ(obj):
        return FAILED
    a = obj.Normal()
    ll, lp = a["LogLikelihoodFunction"], a["LogPriorPDFFunction"]
    lo = np.array([p[1] for p in a["Parameters"]], float)
    hi = np.array([p[2] for p in a["Parameters"]], float)
    d = lo.size

    def f(X):
        X = np.atleast_2d(X)
        inside = np.all((X > lo) & (X < hi), axis=1)
        v = np.full(X.shape[0], -np.inf)
        if inside.any():
            v[inside] = ll(X[inside]) + lp(X[inside])
        v[v <= 0.5 * LOGZERO_HOST] = -np.inf
        return v

    if InitialGuess is not None:
        x = np.asarray(InitialGuess, float).reshape(d)
    elif "Samples" in a:
        S = a["Samples"]
        x = np.asarray(S["Point"][int(np.argmax(S["LogLikelihood"] + S["LogPriorPDF"]))], float)
    else:
        cand = a["_problem"].sample_prior(4096, seed=77, run_id=0)
        x = cand[int(np.argmax(f(cand)))]
    if "Samples" in a and "CrudePosteriorWeight" in a["Samples"]:
        w = a["Samples"]["CrudePosteriorWeight"]
        w = w / w.sum()
        m = w @ a["Samples"]["Point"]
        h = 0.3 * np.sqrt(np.maximum(w @ (a["Samples"]["Point"] - m) ** 2, 1e-300))  # 0.3 posterior sd
    else:
        h = 1e-4 * np.where(np.isfinite(hi - lo), hi - lo, 1.0)
    E = np.eye(d)
    pairs = [(i, j) for i in range(d) for j in range(i + 1, d)]

    def derivatives(x, h):
        pts = [x] + [x + s * h[i] * E[i] for i in range(d) for s in (1, -1)]
        for i, j in pairs:
            pts += [x + si * h[i] * E[i] + sj * h[j] * E[j] for si in (1, -1) for sj in (1, -1)]
        n2 = len(pts)
        pts += [x + s * 0.5 * h[i] * E[i] for i in range(d) for s in (1, -1)]  # half steps: 5-point gradient, O(h^4)
        v = f(np.array(pts))
        f0, fp, fm = v[0], v[1:2 * d + 1:2], v[2:2 * d + 1:2]
        hp, hm = v[n2::2], v[n2 + 1::2]
        g = (8.0 * (hp - hm) - (fp - fm)) / (6.0 * h)
        H = np.diag((fp - 2 * f0 + fm) / h**2)
        for k, (i, j) in enumerate(pairs):
            q = v[2 * d + 1 + 4 * k: 2 * d + 5 + 4 * k]  # (+,+), (+,-), (-,+), (-,-)
            H[i, j] = H[j, i] = (q[0] - q[1] - q[2] + q[3]) / (4 * h[i] * h[j])
        return f0, g, H, bool(np.all(np.isfinite(v)))

    fx = f(x)[0]
    if not np.isfinite(fx):
        return FAILED
    # two passes: stencil at ~0.3 sd of the local Gaussian to get there robustly, then at ~0.05 sd so that the
    # O(h^2) bias of the central differences is far below the Monte Carlo error anything is compared with
    for rel in (0.3, 0.05):
        if rel != 0.3:
            h = h * (rel / 0.3)
        for _ in range(int(MaxIterations)):
            f0, g, H, ok = derivatives(x, h)
            if not ok:
                h = 0.5 * h  # a stencil point left the box / the operator's domain
                continue
            ev, V = np.linalg.eigh(-H)
            if ev.min() > 0:
                h = np.clip(rel / np.sqrt(np.diag(-H)), 0.1 * h, 4.0 * h)
                step = V @ ((V.T @ g) / ev)  # Newton
            else:
                step = g * h * h  # not concave here: scaled gradient ascent
            ts = 2.0 ** -np.arange(0, 12)
            vals = f(x + ts[:, None] * step)  # the whole line search is one batched call
            k = int(np.argmax(vals))
            if not (vals[k] > f0):
                break
            x = x + ts[k] * step
            if vals[k] - f0 < 1e-12 + 8 * np.finfo(float).eps * abs(f0) and ev.min() > 0:
                break
    f0, g, H, ok = derivatives(x, h)
    if not ok:
        return FAILED
    prec = -0.5 * (H + H.T)
    names = a["ParameterSymbols"]
    out = {"LogEvidence": laplaceLogEvidence(f0, prec), "Maximum": (float(f0), dict(zip(names, x.tolist()))),
           "Mean": x, "PrecisionMatrix": prec, "Parameters": names}
    if not np.all(np.linalg.eigvalsh(prec) > 0):
        warnings.warn("approximateEvidence::nonposdef: the Hessian at the maximum is not negative definite")  # LA:214-216
    return out


class GPPrediction(dict):
    """Association input -> MixtureDistribution (GP:355-378), plus the columnar arrays it was built from:
    .points (Q, D), .weights (M,), .means / .sds (M, Q)."""


def predictFromGaussianProcess(obj, pts, _backend_override=None):
    """GP:332-393.  pts: an integer n > 1 (n equally spaced inputs per dimension over the bounds of the data inputs,
    GP:335-342) or a list / matrix of inputs.  Every sample of the run contributes one NormalDistribution per input
    (GP:395-420), mixed with the samples' "CrudePosteriorWeight" (GP:353, 357)."""
    if not inferenceObjectQ(obj):
        return FAILED
    a = obj.Normal()
    if "GaussianProcessData" not in a or "Samples" not in a:
        return FAILED  # the reference's definition does not match and the call stays unevaluated
    S = a["Samples"]
    if "CrudePosteriorWeight" not in S:
        return FAILED  # needs evidenceSampling with PostProcessSamplingRuns > 0 (BS:1237)
    xin = a["Data"][0]
    D = xin.shape[1]
    if isinstance(pts, (int, np.integer)) and not isinstance(pts, bool):
        if pts <= 1 or "Data" not in a:
            return FAILED
        axes = [np.linspace(xin[:, j].min(), xin[:, j].max(), int(pts)) for j in range(D)]  # CoordinateBoundsArray
        grid = np.stack(np.meshgrid(*axes, indexing="ij"), -1).reshape(-1, D)
    else:
        grid = np.asarray(pts, float)
        grid = grid.reshape(-1, 1) if grid.ndim == 1 else grid  # dataNormalForm
        if grid.ndim != 2 or grid.shape[1] != D or not np.all(np.isfinite(grid)):
            return FAILED
    _, first = np.unique(grid, axis=0, return_index=True)  # association keys: a repeated input appears once
    grid = grid[np.sort(first)]
    mean, sd = a["_problem"].gp_predict(S["Point"], grid)
    w = np.asarray(S["CrudePosteriorWeight"], float)
    out = GPPrediction()
    for q in range(grid.shape[0]):
        key = float(grid[q, 0]) if D == 1 else tuple(float(v) for v in grid[q])
        out[key] = MixtureDistribution(w, mean[:, q], sd[:, q])
    out.points, out.weights, out.means, out.sds = grid, w / w.sum(), mean, sd
    return out


def generateStartingPoints(obj, n, seed=1):
    """BS:1046-1068: n i.i.d. draws from the prior, stored under "StartingPoints"."""
    if not inferenceObjectQ(obj):
        return inferenceObject(FAILED)
    a = obj.Normal()
    a["StartingPoints"] = a["_problem"].sample_prior(int(n), seed=seed, run_id=0)
    return inferenceObject(a)


# ----------------------------------------------------------------------------------------------------
NS_DEFAULTS = {  # Options[nestedSampling] BS:837-851 (+ engine knobs)
    "SamplePoolSize": 100, "StartingPoints": "Automatic", "MaxIterations": 10000, "MinIterations": 100,
    "MonteCarloMethod": "Automatic", "MonteCarloSteps": 200, "TerminationFraction": 0.01, "Monitor": True,
    "LogLikelihoodMaximum": "Automatic", "MinMaxAcceptanceRate": (0, 1),
    "PostProcessSamplingRuns": 100, "EmpiricalPosteriorDistributionType": "Simple",
    "BatchSize": 1, "Seed": 1,
}


def _ns_options(opts, allowed):
    bad = [k for k in opts if k not in allowed]
    if bad:
        raise TypeError(f"Unknown option(s) {bad}")  # OptionValue::nodef
    o = dict(NS_DEFAULTS)
    o.update(opts)
    steps = o["MonteCarloSteps"]
    if not isinstance(steps, (int, np.integer)):  # BS:869-878
        warnings.warn(f"nestedSampling::MCSteps: cannot use value {steps!r}; defaulting to 200")
        o["MonteCarloSteps"] = 200
    return o


def _engine_options(be, o, n_runs=1, first_run_id=0):
    lo, hi = o["MinMaxAcceptanceRate"]
    lmax = o.get("LogLikelihoodMaximum", "Automatic")  # a number replaces the running maximum in BS:925-932
    lmax = float(lmax) if isinstance(lmax, (int, float, np.integer, np.floating)) and not isinstance(lmax, bool) else float("nan")
    return be.default_options(loglmax=lmax, pool_size=int(o["SamplePoolSize"]), batch_k=int(o["BatchSize"]),
                              mc_steps=int(o["MonteCarloSteps"]), max_iter=int(o["MaxIterations"]),
                              min_iter=int(o["MinIterations"]), term_frac=float(o["TerminationFraction"]),
                              acc_min=float(lo), acc_max=float(hi), seed=int(o["Seed"]),
                              first_run_id=int(first_run_id), n_runs=int(n_runs))


def _samples_table(s):
    """Columnar form of the per-sample records (BS:907-912, 1009-1015, 825-828)."""
    t = {
        "Point": s["points"], "LogLikelihood": s["logL"], "LogPriorPDF": s["logPrior"],
        "AcceptanceRate": s["acc"],  # NaN = Missing["InitialSample"] (BS:911)
        "PoolSize": s["pool"],
    }
    if s.get("logX") is not None:  # absent for runs fetched only to be merged (combineRuns re-weights, BS:1299)
        t.update({"LogX": s["logX"], "X": np.exp(s["logX"]), "CrudeLogPosteriorWeight": s["crude_logw"]})
    return t


def _result_assoc(s, n):
    M = s["logL"].size
    pts = s["points"]
    return {  # BS:1026-1032
        "Samples": _samples_table(s), "SamplePoolSize": int(n), "GeneratedNestedSamples": int(M - n),
        "TotalSamples": int(M), "ParameterRanges": np.stack([pts.min(0), pts.max(0)], 1),
    }


def nestedSampling(obj, _backend_override=None, **opts):
    """BS:1099-1136.  Options as in the reference, plus "BatchSize" (points replaced per iteration; 1 is the
    reference scheme) and "Seed"."""
    if not inferenceObjectQ(obj):
        return inferenceObject(FAILED)
    o = _ns_options(opts, NS_DEFAULTS)
    a = obj.Normal()
    be = _backend(_backend_override or a.get("_backend"))
    start = o["StartingPoints"]
    if isinstance(start, str):
        start = a.get("StartingPoints")
    if start is not None:
        start = np.asarray(start, dtype=np.float64)
        if start.ndim != 2 or start.shape[1] != len(a["Parameters"]):
            return inferenceObject(FAILED)
        o["SamplePoolSize"] = start.shape[0]
    run = be.RunGroup(a["_problem"], _engine_options(be, o), start)  # "Bad likelihood function" raises (BS:920)
    run.advance(0)
    s = run.fetch(0)
    run.close()
    a.update(_result_assoc(s, o["SamplePoolSize"]))
    if start is not None:
        a["StartingPoints"] = start
    ev = evidenceSampling(a, a["ParameterSymbols"], _backend_override=be,
                          PostProcessSamplingRuns=o["PostProcessSamplingRuns"],
                          EmpiricalPosteriorDistributionType=o["EmpiricalPosteriorDistributionType"], Seed=o["Seed"])
    return inferenceObject(ev)


def _mean_and_error(x, axis=0):  # meanAndError BS:1138-1156: Mean and (n-1) StandardDeviation
    x = np.asarray(x)
    return {"Mean": x.mean(axis), "StandardError": x.std(axis, ddof=1)}


_DERIVED_COLUMNS = ("SampledLogX", "LogPosteriorWeight", "CrudePosteriorWeight", "CrudeLogPosteriorWeight", "X", "LogX")


def evidenceSampling(obj_or_assoc, paramNames=None, _backend_override=None, **opts):
    """BS:1158-1291.  Accepts an inferenceObject (returns one) or an association (returns one)."""
    wrap = isinstance(obj_or_assoc, inferenceObject)
    if wrap and not inferenceObjectQ(obj_or_assoc):
        return inferenceObject(FAILED)
    a = obj_or_assoc.Normal() if wrap else dict(obj_or_assoc)
    o = {"PostProcessSamplingRuns": 100, "EmpiricalPosteriorDistributionType": "Simple", "Seed": 1}
    bad = [k for k in opts if k not in o]
    if bad:
        raise TypeError(f"Unknown option(s) {bad}")
    o.update(opts)
    be = _backend(_backend_override or a.get("_backend"))
    names = paramNames if paramNames is not None else a.get("ParameterSymbols", [])
    S = a["Samples"]
    n = int(a.get("_LiveBlock", a["SamplePoolSize"]))  # see combineRuns "PoolSizes": the tail that acts as the live set
    # calculateWeightsCrude: samples must be sorted by {logL, point} (BS:814) — restore that order first.  Columns a
    # previous evidenceSampling derived (the reference overwrites them in the Join at BS:1239-1251) are dropped and
    # recomputed, so evidenceSampling[obj] on a finished result re-post-processes it (BS:1158-1160).
    S = {k: v for k, v in S.items() if k not in _DERIVED_COLUMNS}
    order = _lex_order(S["Point"], S["LogLikelihood"])
    S = {k: _take(v, order) for k, v in S.items()}
    pool = S.get("PoolSize")
    M = S["LogLikelihood"].size
    if pool is None:
        pool = np.concatenate([np.full(M - n, n), np.arange(n, 0, -1)]).astype(np.int64)
    import time as _time
    _t0 = _time.perf_counter()
    cw = be.crude_weights(S["LogLikelihood"], pool, n)
    _PHASES["crude_weights_s"] = _time.perf_counter() - _t0
    S["LogX"], S["X"], S["CrudeLogPosteriorWeight"] = cw["logX"], np.exp(cw["logX"]), cw["crude_logw"]
    out = dict(a)
    out.update({  # BS:1183-1194
        "CrudeLogEvidence": cw["crude_logZ"], "LogLikelihoodMaximum": cw["logLmax"],
        "LogEstimatedMissingEvidence": cw["log_missing"], "CrudeRelativeEntropy": cw["entropy"],
    })
    nruns = o["PostProcessSamplingRuns"]
    if not (isinstance(nruns, (int, np.integer)) and nruns > 0):  # BS:1195-1197
        out["Samples"] = S
        return inferenceObject(out) if wrap else out
    _t0 = _time.perf_counter()
    ev = be.evidence_sampling(S["Point"], S["LogLikelihood"], pool, n, int(max(nruns, 2)), int(o["Seed"]))
    _PHASES["evidence_sampling_s"] = _time.perf_counter() - _t0
    S["CrudeLogPosteriorWeight"] = S["CrudeLogPosteriorWeight"] - cw["crude_logZ"]          # BS:1236
    S["CrudePosteriorWeight"] = np.exp(S["CrudeLogPosteriorWeight"])                         # BS:1237
    S["SampledLogX"] = {"Mean": ev["slx_mean"], "StandardError": ev["slx_sd"]}                # BS:1244
    S["LogPosteriorWeight"] = {"Mean": ev["logw_mean"], "StandardError": ev["logw_sd"]}       # BS:1245-1250
    srt = _stable_argsort(-S["CrudeLogPosteriorWeight"])                                      # BS:1241
    S = {k: ({kk: vv[srt] for kk, vv in v.items()} if isinstance(v, dict) else v[srt]) for k, v in S.items()}
    _evidence_summary(out, S, ev, names, o)
    return inferenceObject(out) if wrap else out


def _evidence_summary(out, S, ev, names, o):
    """the result keys evidenceSampling derives from the Monte-Carlo draws (BS:1252-1288)"""
    pme = _mean_and_error(ev["pmean"], 0)
    out.update({
        "Samples": S,
        "LogEvidence": {"Mean": float(ev["z"].mean()), "StandardError": float(ev["z"].std(ddof=1))},  # BS:1254
        "ParameterExpectedValues": (  # BS:1255-1262
            {nm: {"Mean": float(pme["Mean"][i]), "StandardError": float(pme["StandardError"][i])}
             for i, nm in enumerate(names)} if len(names) == ev["pmean"].shape[1] else pme),
        "RelativeEntropy": {"Mean": float(ev["H"].mean()), "StandardError": float(ev["H"].std(ddof=1))},  # BS:1263
        "EmpiricalPosteriorDistribution": {"Type": o["EmpiricalPosteriorDistributionType"],  # BS:1269-1288
                                           "Weights": S["CrudePosteriorWeight"], "Points": S["Point"]},
    })


def _stable_argsort(key):
    """np.argsort(key, kind="stable") for keys with few ties: the default (vectorised) sort, then only the runs of equal
    keys are put back into index order — SortBy is stable (BS:1241), numpy's stable sort is ~8x slower on 3e5 floats."""
    o = np.argsort(key)
    sk = key[o]
    tie = sk[1:] == sk[:-1]
    if tie.any():
        m = np.zeros(o.size, dtype=bool)
        m[1:] |= tie
        m[:-1] |= tie
        sub = o[m]
        o[m] = sub[np.lexsort((sub, key[sub]))]
    return o


class _Identity:
    """Stands for the identity permutation: _take() then skips the gather (a copy of every column of a list with
    hundreds of thousands of samples is most of the host time of a merge)."""

    size = None


_IDENTITY = _Identity()


def _take(v, order):
    return v if order is _IDENTITY else v[order]


def _lex_order(pts, logL=None):
    """Stable order by (logL, point) — SortBy[{#LogLikelihood, #Point}&] (BS:814) — or by point alone.
    One stable argsort on the leading key (timsort: lists that arrive sorted, or as a concatenation of sorted runs,
    cost O(M) / O(M log R)); only the (rare) runs of equal leading keys are refined with a lexsort."""
    lead = logL if logL is not None else pts[:, 0]
    if lead.size > 1 and np.all(lead[1:] > lead[:-1]):
        return _IDENTITY  # strictly increasing already (the engine's fetch order, a merged list): nothing to move
    o = np.argsort(lead, kind="stable")
    sl = lead[o]
    tie = sl[1:] == sl[:-1]
    if tie.any():
        m = np.zeros(o.size, dtype=bool)
        m[1:] |= tie
        m[:-1] |= tie
        sub = o[m]
        # stable within fully equal rows: the original position is the last key
        keys = (sub,) + tuple(pts[sub, j] for j in range(pts.shape[1] - 1, -1, -1)) + (lead[sub],)
        o[m] = sub[np.lexsort(keys)]
    return o


def _merge_samples(tables, pool_sizes):
    """combineRuns BS:1293-1297: Join, DeleteDuplicatesBy Point (first kept), SortBy {logL, Point};
    per-sample pool size = sum over runs of that run's pool size at the sample's likelihood level.
    O(M log R) for M samples in R runs: ONE stable sort of the joined list (a concatenation of sorted runs) serves
    the merge, the duplicate removal and the pool sizes.  Duplicates are a walk's unmoved copies of a live point —
    same point, same likelihood — so after the stable sort they are adjacent with the Join-order first in front
    (DeleteDuplicatesBy keeps exactly that one).  Every run's pool size is a step function of the likelihood level
    that changes at its own samples, so the sum over runs is one cumulative sum over the joined samples in level
    order (the reference recomputes X from scratch instead, calculateXValues BS:785-799)."""
    per = []
    base = 0
    for t, n in zip(tables, pool_sizes):
        o = _lex_order(t["Point"], t["LogLikelihood"])
        tl = _take(t["LogLikelihood"], o)
        tp = t.get("PoolSize")
        tp = _take(tp, o) if tp is not None else np.concatenate([np.full(max(tl.size - n, 0), n), np.arange(min(n, tl.size), 0, -1)])
        tp = np.asarray(tp, dtype=np.int64)
        if tl.size:
            base += int(tp[0])
        per.append((o, tl, np.diff(np.concatenate([tp, [0]])) if tl.size else tp))
    # Join in run order; every run's rows travel in that run's sorted order, `pos` remembers the Join position
    pts = np.concatenate([_take(t["Point"], o) for t, (o, _, _) in zip(tables, per)])
    cols = {k: np.concatenate([_take(t[k], o) for t, (o, _, _) in zip(tables, per)]) for k in ("LogLikelihood", "LogPriorPDF", "AcceptanceRate")}
    sizes = [tl.size for _, tl, _ in per]
    rid = np.repeat(np.arange(len(per)), sizes)
    offs = np.cumsum([0] + sizes[:-1])
    pos = np.concatenate([(np.arange(sz) if o is _IDENTITY else o) + off for (o, _, _), off, sz in zip(per, offs, sizes)])
    delta = np.concatenate([d for _, _, d in per])
    order = _lex_order(pts, cols["LogLikelihood"])
    pts, rid, delta, pos = _take(pts, order), _take(rid, order), _take(delta, order), _take(pos, order)
    cols = {k: _take(v, order) for k, v in cols.items()}
    L = cols["LogLikelihood"]
    # pool size at a sample = base + sum of the deltas of all run samples STRICTLY below its level
    csum = np.concatenate([[0], np.cumsum(delta)])
    new_level = np.concatenate([[True], L[1:] != L[:-1]])
    first_of_level = np.flatnonzero(new_level)
    pool = base + csum[first_of_level[np.cumsum(new_level) - 1]]
    # DeleteDuplicatesBy[Point] (first in Join order kept).  Duplicates are rare (a walk with no accepted move returns
    # a copy of a live point): rows are hashed, the hashes sorted (integer sort), and only colliding rows are compared
    # exactly.
    keep = _first_occurrence_mask(pts, pos)
    if not keep.all():
        pts, rid, pool = pts[keep], rid[keep], pool[keep]
        cols = {k: v[keep] for k, v in cols.items()}
    out = {"Point": pts, "PoolSize": pool.astype(np.int64), "RunIndex": rid}
    out.update(cols)
    return out


def _first_occurrence_mask(pts, pos):
    """True for the rows of `pts` that are the first (smallest `pos`) among rows with equal values."""
    M, d = pts.shape
    keep = np.ones(M, dtype=bool)
    if M < 2:
        return keep
    bits = np.ascontiguousarray(pts + 0.0).view(np.uint64).reshape(M, d)  # + 0.0: -0.0 and 0.0 are the same point
    mult = (np.arange(d, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(0xD6E8FEB86659FD93)) | np.uint64(1)
    with np.errstate(over="ignore"):
        h = np.bitwise_xor.reduce((bits ^ (bits >> np.uint64(29))) * mult, axis=1)
    hs = np.sort(h)
    coll = hs[1:][hs[1:] == hs[:-1]]
    if coll.size == 0:
        return keep
    cand = np.flatnonzero(np.isin(h, np.unique(coll)))
    sub = pts[cand]
    o = np.lexsort((pos[cand],) + tuple(sub[:, j] for j in range(d - 1, -1, -1)))
    so = sub[o]
    dup = np.concatenate([[False], np.all(so[1:] == so[:-1], axis=1)])
    keep[cand[o[dup]]] = False
    return keep


def _reference_pool_structure(t, n):
    """True when a run's pool-size column is the reference's: constant n for the deleted points, n..1 for the
    final live set (BS:785-799) — i.e. the run replaced one point per iteration."""
    tp = t.get("PoolSize")
    if tp is None:
        return True
    o = _lex_order(t["Point"], t["LogLikelihood"])
    tp = _take(np.asarray(tp), o)
    M = tp.size
    return bool(M >= n and np.all(tp[:M - n] == n) and np.array_equal(tp[M - n:], np.arange(n, 0, -1)))


def _merged_assoc(res, n_tot, scheme):
    """result keys of a device combine (engine.combine_runs / RunGroup.combine): BS:1307-1309, BS:1183-1194"""
    M = res["Samples"]["LogLikelihood"].size
    out = {"SamplePoolSize": n_tot, "GeneratedNestedSamples": M - n_tot, "TotalSamples": M, "MergeScheme": scheme,
           "CrudeLogEvidence": res["crude_logZ"], "LogLikelihoodMaximum": res["logLmax"],
           "LogEstimatedMissingEvidence": res["log_missing"], "CrudeRelativeEntropy": res["entropy"]}
    if scheme != "Reference":
        out["_LiveBlock"] = int(res["n_live"])
    return out


def _sorted_run_table(t, n):
    """one run's columns for the device merge: sorted by {logL, point} (the engine's fetch order: nothing moves) and
    with the run's own pool sizes (the reference's structure BS:785-799 when the run does not carry them)"""
    o = _lex_order(t["Point"], t["LogLikelihood"])
    out = {k: _take(t[k], o) for k in ("Point", "LogLikelihood", "LogPriorPDF", "AcceptanceRate") if k in t}
    tp = t.get("PoolSize")
    M = out["LogLikelihood"].size
    out["PoolSize"] = (_take(np.asarray(tp), o) if tp is not None
                       else np.concatenate([np.full(max(M - n, 0), n), np.arange(min(n, M), 0, -1)]))
    return out


_HOST_MERGE = False  # tests: force the numpy merge on a backend that has the device one
_HOST_GATHER = False  # tests: gather the per-GPU merges through the host (all_gather_object) under an NCCL process group
_PHASES = {}  # wall-clock seconds of the last combineRuns / evidenceSampling call, by phase (diagnostics for bench.py)


def combineRuns(*results, _backend_override=None, **opts):
    """BS:1293-1315.  The merged list is re-weighted as ONE run (evidenceSampling -> calculateXValues BS:785-799).
    "MergeScheme" (not in the reference) selects the X sequence of the merged list:
      "Reference"  the literal formula: pool size Total[SamplePoolSize] for the first M - n_tot samples, then the
                   n_tot best as a live set n_tot..1 (BS:1307-1309 feeding BS:785-799);
      "PoolSizes"  the pool size at every sample is the SUM over runs of that run's pool size at the sample's
                   likelihood level (the merge rule of dynamic nested sampling), used consistently to the last sample;
                   the only consistent choice when runs replaced K > 1 points per iteration ("BatchSize"), whose
                   own pool sizes are n, n-1, ..., n-K+1 and not a constant;
      "Automatic"  (default) "Reference" when every run has the reference's pool structure, else "PoolSizes"."""
    if len(results) < 1 or not all(inferenceObjectQ(r) for r in results):
        return inferenceObject(FAILED)
    scheme = opts.pop("MergeScheme", "Automatic")
    if scheme not in ("Automatic", "Reference", "PoolSizes"):
        raise TypeError(f"Unknown MergeScheme {scheme!r}")
    assocs = [r.Normal() for r in results]
    pools = [int(a["SamplePoolSize"]) for a in assocs]
    if scheme == "Automatic":
        scheme = "Reference" if all(_reference_pool_structure(a["Samples"], n) for a, n in zip(assocs, pools)) else "PoolSizes"
    import time as _time
    _PHASES.clear()
    n_tot = int(sum(pools))
    be = _backend(_backend_override or assocs[0].get("_backend"))
    ev_opts = {"PostProcessSamplingRuns": 100, "EmpiricalPosteriorDistributionType": "Simple", "Seed": 1}
    bad = [k for k in opts if k not in ev_opts]
    if bad:
        raise TypeError(f"Unknown option(s) {bad}")
    ev_opts.update(opts)
    nruns = ev_opts["PostProcessSamplingRuns"]
    if hasattr(be, "combine_runs") and isinstance(nruns, (int, np.integer)) and nruns > 0 and not _HOST_MERGE:
        # the whole of combineRuns -> evidenceSampling on the device (csrc/merge.cu): one call, one table back
        _t0 = _time.perf_counter()
        res = be.combine_runs([_sorted_run_table(x["Samples"], n) for x, n in zip(assocs, pools)], scheme == "Reference",
                              n_tot, int(max(nruns, 2)), int(ev_opts["Seed"]))
        _PHASES["device_combine_s"] = _time.perf_counter() - _t0
        a = dict(assocs[0])
        a.pop("_LiveBlock", None)
        a.update(_merged_assoc(res, n_tot, scheme))
        _evidence_summary(a, res["Samples"], res, a.get("ParameterSymbols", []), ev_opts)
        return inferenceObject(a)
    _t0 = _time.perf_counter()
    merged = _merge_samples([a["Samples"] for a in assocs], pools)
    _PHASES["merge_s"] = _time.perf_counter() - _t0
    M = merged["LogLikelihood"].size
    a = dict(assocs[0])
    a.pop("_LiveBlock", None)
    a.update({
        "Samples": merged,
        "LogLikelihoodMaximum": max(float(np.max(x["Samples"]["LogLikelihood"])) for x in assocs),  # BS:1306
        "SamplePoolSize": n_tot, "GeneratedNestedSamples": M - n_tot, "TotalSamples": M,          # BS:1307-1309
        "MergeScheme": scheme,
    })
    if scheme == "Reference":
        merged["PoolSize"] = np.concatenate([np.full(max(M - n_tot, 0), n_tot), np.arange(min(n_tot, M), 0, -1)]).astype(np.int64)
    else:
        # summed pool sizes to the end.  Once every run is inside its final live set the sum falls by one per sample:
        # that tail (length >= the largest run pool) IS a live set in the sense of BS:791-797, and is handed to
        # calculateXValues / the order-statistics draws of BS:1209-1217 as such.
        pool = merged["PoolSize"]
        tail = np.arange(M, 0, -1)
        agree = pool == tail
        live = int(M - (np.flatnonzero(~agree)[-1] + 1)) if not agree.all() else M
        a["_LiveBlock"] = max(live, 1)
    return inferenceObject(evidenceSampling(a, a.get("ParameterSymbols"), _backend_override=_backend_override, **opts))


def _shard(n_runs, rank, world):
    """contiguous block of run ids for this rank"""
    base, rem = divmod(n_runs, world)
    lo = rank * base + min(rank, rem)
    return lo, base + (1 if rank < rem else 0)


def _fetch_for_merge(grp, i):
    try:
        return grp.fetch(i, weights=False)  # the merged list is re-weighted as one run: per-run weights are not needed
    except TypeError:
        return grp.fetch(i)


def parallelNestedSampling(obj, _backend_override=None, **opts):
    """BS:1317-1371.  "ParallelRuns" independent runs, each drawing its own starting points, advanced in lock
    step on the GPU (one library call instead of ParallelTable over subkernels) and merged with combineRuns.
    Under torch.distributed (one process per GPU) the runs are sharded across ranks with no data-path
    collective; the per-run sample lists are gathered and every rank returns the merged object."""
    if not inferenceObjectQ(obj):
        return inferenceObject(FAILED)
    allowed = dict(NS_DEFAULTS)
    allowed["ParallelRuns"] = 4  # BS:1369
    o = _ns_options(opts, allowed)
    R = int(o.get("ParallelRuns", 4))
    a = obj.Normal()
    if a.get("StartingPoints") is not None:  # BS:1320-1332
        n0 = len(a["StartingPoints"])
        warnings.warn("parallelNestedSampling::startingPts: pre-specified starting points are ignored; "
                      f'continuing with "SamplePoolSize" -> {n0}')
        o["SamplePoolSize"] = n0
        a.pop("StartingPoints")
    be = _backend(_backend_override or a.get("_backend"))
    rank, world = 0, 1
    dist = None
    try:
        import torch.distributed as dist_mod
        if dist_mod.is_available() and dist_mod.is_initialized():
            dist = dist_mod
            rank, world = dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    import time as _time
    tm = {}
    t0 = _time.perf_counter()
    first, count = _shard(R, rank, world)
    n = int(o["SamplePoolSize"])
    nruns = o["PostProcessSamplingRuns"]
    ev_opts = {k: o[k] for k in ("PostProcessSamplingRuns", "EmpiricalPosteriorDistributionType", "Seed")}
    # device merge (csrc/merge.cu): the runs never travel to the host one by one.  One GPU: the whole of combineRuns ->
    # evidenceSampling runs on the group's device state; several GPUs: every rank merges its own runs on its GPU, the
    # per-rank merges are gathered, and merging those equals merging all runs at once (summed pool sizes are additive).
    on_device = (hasattr(be, "combine_runs") and hasattr(be.RunGroup, "combine") and isinstance(nruns, (int, np.integer))
                 and nruns > 0 and not _HOST_MERGE)
    nccl_gather = (on_device and world > 1 and hasattr(be, "combine_runs_dev") and dist.get_backend() == "nccl"
                   and not _HOST_GATHER)
    reference_scheme = int(o.get("BatchSize", 1)) == 1  # K = 1 runs have the reference's pool structure (combineRuns "Automatic")
    local = []
    if on_device:
        _PHASES.clear()
        res, part = None, None
        if count > 0:
            grp = be.RunGroup(a["_problem"], _engine_options(be, o, n_runs=count, first_run_id=first), None)
            grp.advance(0)
            tm["device_loop_s"] = _time.perf_counter() - t0
            t1 = _time.perf_counter()
            if world == 1:
                res = grp.combine(reference_scheme, int(max(nruns, 2)), int(o["Seed"]))
            elif nccl_gather:
                part = grp.merge_dev()[0]
            else:
                part = grp.merge()[0]
            grp.close()
            tm["device_merge_s"] = _time.perf_counter() - t1
        if world > 1 and nccl_gather:
            # the per-GPU merges stay in device memory and are all-gathered over NCCL (the one exchange step of the
            # call); every rank then merges the gathered lists on its GPU
            import torch
            t1 = _time.perf_counter()
            width = len(a["Parameters"]) + 5 if part is None else part.shape[1]
            mine = part if part is not None else torch.empty((0, width), dtype=torch.float64, device="cuda")
            cnt = torch.tensor([mine.shape[0]], dtype=torch.int64, device="cuda")
            cnts = torch.empty(world, dtype=torch.int64, device="cuda")
            dist.all_gather_into_tensor(cnts, cnt)
            cnts = cnts.tolist()
            padded = torch.zeros((max(cnts), width), dtype=torch.float64, device="cuda")
            padded[:mine.shape[0]] = mine
            allt = torch.empty((world, max(cnts), width), dtype=torch.float64, device="cuda")
            dist.all_gather_into_tensor(allt.view(-1), padded.view(-1))
            tm["gather_s"] = _time.perf_counter() - t1
            t1 = _time.perf_counter()
            res = be.combine_runs_dev([allt[r, :cnts[r]] for r in range(world)], reference_scheme, R * n, int(max(nruns, 2)),
                                      int(o["Seed"]))
            tm["device_combine_s"] = _time.perf_counter() - t1
        elif world > 1:
            t1 = _time.perf_counter()
            gathered = [None] * world
            dist.all_gather_object(gathered, part)  # host gather of the per-GPU merges (no NCCL process group)
            tm["gather_s"] = _time.perf_counter() - t1
            t1 = _time.perf_counter()
            res = be.combine_runs([g for g in gathered if g is not None], reference_scheme, R * n, int(max(nruns, 2)),
                                  int(o["Seed"]))
            tm["device_combine_s"] = _time.perf_counter() - t1
        t1 = _time.perf_counter()
        out = dict(a)
        out.update(_merged_assoc(res, R * n, "Reference" if reference_scheme else "PoolSizes"))
        p0 = res["Samples"]["Point"][res["Samples"]["RunIndex"] == 0]  # combineRuns keeps First[results] (BS:1299-1305)
        out["ParameterRanges"] = np.stack([p0.min(0), p0.max(0)], 1)
        _evidence_summary(out, res["Samples"], res, out.get("ParameterSymbols", []), ev_opts)
        tm["assemble_s"] = _time.perf_counter() - t1
        tm["total_s"] = _time.perf_counter() - t0
        out["_Timing"] = tm
        return inferenceObject(out)
    if count > 0:
        grp = be.RunGroup(a["_problem"], _engine_options(be, o, n_runs=count, first_run_id=first), None)
        grp.advance(0)
        tm["device_loop_s"] = _time.perf_counter() - t0
        t1 = _time.perf_counter()
        local = [(first + i, _fetch_for_merge(grp, i)) for i in range(count)]
        grp.close()
        tm["fetch_s"] = _time.perf_counter() - t1
    t1 = _time.perf_counter()
    if dist is not None and world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, local)  # host merge only: no collective on the data path
        local = [x for part in gathered for x in part]
    tm["gather_s"] = _time.perf_counter() - t1
    local.sort(key=lambda t: t[0])
    t1 = _time.perf_counter()
    runs = []
    for _, s in local:
        ra = dict(a)
        ra.update(_result_assoc(s, n))
        runs.append(inferenceObject(ra))
    res = combineRuns(*runs, _backend_override=be, PostProcessSamplingRuns=o["PostProcessSamplingRuns"],
                      EmpiricalPosteriorDistributionType=o["EmpiricalPosteriorDistributionType"], Seed=o["Seed"])
    tm["combine_and_evidence_s"] = _time.perf_counter() - t1
    tm.update({"combine_" + k: v for k, v in _PHASES.items()})
    tm["total_s"] = _time.perf_counter() - t0
    if inferenceObjectQ(res):
        res._assoc["_Timing"] = tm  # wall-clock phases of this call on this rank (bench.py; keys() hides "_" entries)
    return res
