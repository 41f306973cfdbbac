"""The LibraryLink shim cannot run here (no Wolfram Engine), but it must at least compile and link against
libbinest.so, export the entry points the WL package loads, and contain no arithmetic of its own."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "bayesianinference_b200", "wl", "librarylink_shim.c")
WL = os.path.join(ROOT, "bayesianinference_b200", "wl", "BayesianInferenceB200.wl")


def test_shim_compiles_links_and_exports(tmp_path):
    out = tmp_path / "binestLink.so"
    cmd = ["/usr/bin/gcc", "-shared", "-fPIC", "-Wall", "-Werror=implicit-function-declaration",
           "-I", os.path.join(ROOT, "tests", "wl_stub"), "-I", os.path.join(ROOT, "include"), SHIM, "-o", str(out),
           "-L", os.path.join(ROOT, "bayesianinference_b200"), "-lbinest"]
    subprocess.check_call(cmd)
    syms = subprocess.check_output(["nm", "-D", "--defined-only", str(out)], text=True)
    loaded = set(re.findall(r'll\["(binest[A-Za-z]+)"', open(WL).read()))
    assert len(loaded) >= 12
    for name in loaded | {"WolframLibrary_getVersion", "WolframLibrary_initialize", "WolframLibrary_uninitialize"}:
        assert re.search(rf"\b{name}\b", syms), f"{name} loaded by the WL package but not exported by the shim"


def test_wl_package_keeps_the_reference_api_surface():
    src = open(WL).read()
    for sym in ("defineInferenceProblem", "nestedSampling", "parallelNestedSampling", "evidenceSampling", "combineRuns",
                "generateStartingPoints", "inferenceObject", "defineGaussianProcess", "predictFromGaussianProcess",
                "predictiveDistribution", "createMCMCChain", "iterateMCMC", "approximateEvidence", "laplaceLogEvidence"):
        assert re.search(rf"^{sym}::usage", src, flags=re.M), sym
    for opt in ('"SamplePoolSize" -> 100', '"MaxIterations" -> 10000', '"MinIterations" -> 100', '"MonteCarloSteps" -> 200',
                '"TerminationFraction" -> 0.01', '"MinMaxAcceptanceRate" -> {0, 1}', '"PostProcessSamplingRuns" -> 100',
                '"ParallelRuns" :> 4', '"CovarianceLearnDelay" -> 20', '"InitialCovariance" -> 1'):
        assert opt in src, opt  # Options of BS:833-851, 1366-1371


def test_wl_package_brackets_balance():
    """The package cannot be parsed here (no Wolfram Engine); at least every bracket, association delimiter, string and
    comment must close — the class of slip a kernel would report first."""
    src = open(WL).read()
    pairs = {")": "(", "]": "[", "}": "{"}
    stack, i, line, depth, in_str = [], 0, 1, 0, False
    while i < len(src):
        c = src[i]
        line += c == "\n"
        if in_str:
            if c == "\\":
                i += 2
                continue
            in_str = c != '"'
        elif depth:
            if src.startswith("(*", i):
                depth += 1
                i += 1
            elif src.startswith("*)", i):
                depth -= 1
                i += 1
        elif src.startswith("(*", i):
            depth = 1
            i += 1
        elif c == '"':
            in_str = True
        elif src.startswith("<|", i):
            stack.append(("<|", line))
            i += 1
        elif src.startswith("|>", i):
            assert stack and stack[-1][0] == "<|", f"unmatched |> at line {line}"
            stack.pop()
            i += 1
        elif c in "([{":
            stack.append((c, line))
        elif c in ")]}":
            assert stack and stack[-1][0] == pairs[c], f"unmatched {c} at line {line} (open: {stack[-1:]})"
            stack.pop()
        i += 1
    assert not stack and not in_str and depth == 0, (stack[:3], in_str, depth)
