(* ::Package:: *)

(* BayesianInferenceB200` — Wolfram Language host for the B200-native nested-sampling engine.

   Drop-in for the nested-sampling path of ssmit1986/BayesianInference: the same public symbols with the same
   options and result keys; everything numerical happens in libbinest.so (hand-written sm_100a CUDA), reached only
   through the LibraryLink shim librarylink_shim.c.  This file is the reference's own host language; it cannot be
   executed in the build image (no Wolfram Engine), so all logic that can live behind the C ABI does, and the Python
   mirror bayesianinference_b200/api.py (same structure, tested) documents the intended behaviour line by line.

   Reference locations (BayesianInference/Kernel/):
     defineInferenceProblem   BayesianStatistics.wl:148-308     nestedSampling          :1099-1136
     nestedSamplingInternal   BayesianStatistics.wl:859-1040    parallelNestedSampling  :1317-1371
     evidenceSampling         BayesianStatistics.wl:1158-1291   combineRuns             :1293-1315
     generateStartingPoints   BayesianStatistics.wl:1042-1097   inferenceObject         BayesianUtilities.wl:107-138
     defineGaussianProcess    BayesianGaussianProcess.wl:201-330  predictFromGaussianProcess :332-422
     predictiveDistribution   BayesianStatistics.wl:1373-1483     createMCMCChain / iterateMCMC :630-703
*)

BeginPackage["BayesianInferenceB200`"];

defineInferenceProblem::usage = "defineInferenceProblem[rules...] — as in BayesianInference`; \"GeneratingDistribution\" must match the GPU operator table.";
nestedSampling::usage = "nestedSampling[inferenceObject, opts] runs nested sampling on the GPU.";
parallelNestedSampling::usage = "parallelNestedSampling[inferenceObject, opts] runs \"ParallelRuns\" independent runs in one library call and merges them.";
evidenceSampling::usage = "evidenceSampling[inferenceObject, opts] estimates the evidence error by Monte-Carlo draws of the X sequence.";
combineRuns::usage = "combineRuns[obj1, obj2, ...] merges nested-sampling runs.";
generateStartingPoints::usage = "generateStartingPoints[inferenceObject, n] draws n points from the prior.";
inferenceObject::usage = "inferenceObject[assoc] wraps the results; obj[\"Key\"] extracts a property.";
inferenceObjectQ::usage = "inferenceObjectQ[obj]";
defineGaussianProcess::usage = "defineGaussianProcess[in -> out, squaredExponentialKernel[sf, ell], nuggetVariance[sn], None, {{sf, lo, hi}, {ell, lo, hi}, {sn, lo, hi}}, prior] — GP regression with the squared-exponential kernel sf^2 Exp[-|x - x'|^2/(2 ell^2)] and nugget sn^2 (the GPU operator).";
predictFromGaussianProcess::usage = "predictFromGaussianProcess[inferenceObject, pts] gives, for every input, the MixtureDistribution of the samples' Gaussian predictives weighted by CrudePosteriorWeight.";
predictiveDistribution::usage = "predictiveDistribution[inferenceObject] or predictiveDistribution[inferenceObject, inputs] gives the posterior predictive as a MixtureDistribution over the samples (per input for regression problems); a trailing \"MaximumLikelihood\" or \"MAP\" uses the single best sample.";
createMCMCChain::usage = "createMCMCChain[inferenceObject, startPt, opts] creates an adaptive-Metropolis chain on the log posterior (GPU operators); options \"InitialCovariance\", \"CovarianceLearnDelay\", \"Seed\".";
iterateMCMC::usage = "iterateMCMC[chain, n] advances the chain and returns the n visited states; iterateMCMC[chain, {n, thin}] keeps one state every thin steps.";
laplaceLogEvidence::usage = "laplaceLogEvidence[max, precisionMatrix] = max + (n Log[2 Pi] - Log[Det[precisionMatrix]])/2.";
approximateEvidence::usage = "approximateEvidence[inferenceObject] maximises the log posterior of the GPU operators and returns the Laplace evidence, the maximum and the precision matrix.";
squaredExponentialKernel::usage = "squaredExponentialKernel[sf, ell] — kernel descriptor for defineGaussianProcess.";
nuggetVariance::usage = "nuggetVariance[sn] — nugget sn^2 descriptor for defineGaussianProcess.";
$binestLibrary::usage = "Path of the compiled LibraryLink shim (binestLink).";
categoricalSoftmax::usage = "categoricalSoftmax[{{w11,..,w1F,b1},...}, {x1,..,xF}] — softmax classification with reference class K.";

Begin["`Private`"];

$MachineLogZero = -$MaxMachineNumber; (* BU:47: -Statistics`Library`MachineInfinity *)

(* ------------------------------------------------------------------ LibraryLink bindings: no logic *)
$binestLibrary = FindLibrary["binestLink"];
ll[name_, args_, res_] := ll[name, args, res] = LibraryFunctionLoad[$binestLibrary, name, args, res];
check[rc_Integer] /; rc =!= 0 := (Message[inferenceObject::binest, ll["binestLastError", {}, "UTF8String"][]]; $Failed);
check[other_] := other;
inferenceObject::binest = "libbinest: `1`";

binestInit := ll["binestInit", {Real, Integer}, Integer];
binestProblemCreate := ll["binestProblemCreate", {Integer, {Integer, 1}, {Real, 2, "Constant"}, {Real, 2, "Constant"},
    {Integer, 1}, {Real, 1}, {Real, 1}, {Real, 1}, {Real, 1}}, Integer];
binestLogLike := ll["binestLogLike", {Integer, {Real, 2, "Constant"}}, {Real, 1}];
binestLogPrior := ll["binestLogPrior", {Integer, {Real, 2, "Constant"}}, {Real, 1}];
binestChainCreate := ll["binestChainCreate", {Integer, {Real, 2, "Constant"}, {Real, 2, "Constant"}, Integer, Integer}, Integer];
binestChainIterate := ll["binestChainIterate", {Integer, Integer, Integer, Integer}, {Real, 3}];
binestChainState := ll["binestChainState", {Integer, Integer, Integer}, {Real, 1}];
binestChainFree := ll["binestChainFree", {Integer}, Integer];
binestPredictiveComponents := ll["binestPredictiveComponents", {Integer, {Real, 2, "Constant"}, {Real, 2, "Constant"}}, {Real, 3}];
binestGPPredict := ll["binestGPPredict", {Integer, {Real, 2, "Constant"}, {Real, 2, "Constant"}}, {Real, 3}];
binestSamplePrior := ll["binestSamplePrior", {Integer, Integer, Integer, Integer}, {Real, 2}];
binestRunCreate := ll["binestRunCreate", {Integer, {Integer, 1}, {Real, 1}, {Real, _, "Constant"}}, Integer];
binestRunAdvance := ll["binestRunAdvance", {Integer, Integer}, Integer];
binestRunFetch := ll["binestRunFetch", {Integer, Integer, Integer}, {Real, 2}];
binestRunFree := ll["binestRunFree", {Integer}, Integer];
binestRunCombine := ll["binestRunCombine", {Integer, Integer, Integer, Integer, Integer}, {Real, 2}];
binestEvidenceSampling := ll["binestEvidenceSampling", {{Real, 2, "Constant"}, {Real, 1, "Constant"}, {Integer, 1, "Constant"}, Integer, Integer, Integer}, {Real, 1}];
binestCrudeWeights := ll["binestCrudeWeights", {{Real, 1, "Constant"}, {Integer, 1, "Constant"}, Integer}, {Real, 1}];

initialised = False;
ensureInit[] := If[!initialised, check @ binestInit[$MachineLogZero, -1]; initialised = True];

(* ------------------------------------------------------------------ inferenceObject (BU:107-138) *)
inferenceObject[assoc_?AssociationQ][prop_] := assoc[prop];
inferenceObject /: Normal[inferenceObject[assoc_]] := assoc;
inferenceObjectQ[inferenceObject[_?AssociationQ]] := True;
inferenceObjectQ[___] := False;

paramSpecPattern = {_Symbol, _?NumericQ | DirectedInfinity[-1], _?NumericQ | DirectedInfinity[1]};

(* ------------------------------------------------------------------ operator table (SURVEY.md Appendix A) *)
(* returns {opId, iparam, inputs, outputs} or $Failed *)
operatorFromDistribution[NormalDistribution[mu_Symbol, sigma_Symbol], data_?VectorQ, {mu_, sigma_}, _] :=
    {1, {0, 0, 0, 0}, List /@ N[data], {{}}};
operatorFromDistribution[NormalDistribution[poly_, sigma_Symbol], Rule[in_, out_], params_List, {x_Symbol}] /;
        PolynomialQ[poly, x] && Most[params] === CoefficientList[poly, x] && Last[params] === sigma :=
    {2, {Exponent[poly, x], 0, 0, 0}, ArrayReshape[N[in], {Length[in], 1}], ArrayReshape[N[out], {Length[out], 1}]};
operatorFromDistribution[categoricalSoftmax[blocks_?MatrixQ, vars_List], Rule[in_?MatrixQ, labels_], params_List, vars_List] /;
        Flatten[blocks] === params :=
    {3, {0, Length[blocks] + 1, 0, 0}, N[in], ArrayReshape[N[labels], {Length[labels], 1}]};
operatorFromDistribution[GeometricBrownianMotionProcess[mu_Symbol, sigma_Symbol, _], ts_TemporalData, {mu_, sigma_}, _] :=
    {4, {0, 0, 0, 0}, List /@ N[ts["Times"]], List /@ N[ts["Values"]]}; (* TemporalData adaptor BS:511-515 *)
operatorFromDistribution[gpOperator[sf_Symbol, ell_Symbol, sn_Symbol], Rule[in_?MatrixQ, out_?MatrixQ], {sf_, ell_, sn_}, _] :=
    {5, {0, 0, Last[Dimensions[in]], 0}, N[in], N[out]}; (* GP marginal likelihood, GP:27-61, 130-199 *)
operatorFromDistribution[dist_, ___] := (Message[defineInferenceProblem::logLike, dist]; $Failed);
defineInferenceProblem::logLike = "`1` is not in the GPU operator table; there is no CPU fallback."; (* cf. BS:456-459 *)
defineInferenceProblem::insuffInfo = "Not enough information was provided to define the problem"; (* BS:148-152 *)

priorKinds[prior_List, params_] := MapThread[
    Function[{spec, par},
        Switch[spec,
            "LocationParameter" | _UniformDistribution, {1, 0., 1.},   (* BS:37-39 *)
            "ScaleParameter", {2, 0., 1.},                              (* BS:42-48 *)
            NormalDistribution[_?NumericQ, _?NumericQ], {3, N[spec[[1]]], N[spec[[2]]]}, (* BS:51-59 *)
            _, Throw[$Failed, "problemDef"]]],
    {prior, params}];

defineInferenceProblem[rules__Rule] := defineInferenceProblem[Association[rules]];
defineInferenceProblem[assoc_?AssociationQ] := Catch[
    Module[{params, names, kinds, op, handle, test},
        If[!AllTrue[{"Data", "Parameters", "GeneratingDistribution", "PriorDistribution"}, KeyExistsQ[assoc, #] &],
            Message[defineInferenceProblem::insuffInfo]; Throw[$Failed, "problemDef"]];
        ensureInit[];
        params = Replace[assoc["Parameters"], s_Symbol :> {s, -Infinity, Infinity}, {1}]; (* paramNormalForm BS:133-145 *)
        names = params[[All, 1]];
        kinds = priorKinds[assoc["PriorDistribution"], params];
        op = operatorFromDistribution[assoc["GeneratingDistribution"], assoc["Data"], names,
            Lookup[assoc, "IndependentVariables", {}]];
        If[op === $Failed, Throw[$Failed, "problemDef"]];
        handle = check @ binestProblemCreate[op[[1]], op[[2]], op[[3]], op[[4]], kinds[[All, 1]],
            N[params[[All, 2]]], N[params[[All, 3]]], kinds[[All, 2]], kinds[[All, 3]]];
        If[handle === $Failed, Throw[$Failed, "problemDef"]];
        (* BS:276-298: both functions must be numeric on random points of the box *)
        test = binestSamplePrior[handle, 100, 20260, 0];
        If[!VectorQ[binestLogLike[handle, test], NumericQ] || !VectorQ[binestLogPrior[handle, test], NumericQ],
            Throw[$Failed, "problemDef"]];
        inferenceObject @ Join[assoc, <|
            "Parameters" -> params, "ParameterSymbols" -> names,
            "LogLikelihoodFunction" -> Function[pts, binestLogLike[handle, If[VectorQ[pts], {pts}, pts]]],  (* Listable, BS:499 *)
            "LogPriorPDFFunction" -> Function[pts, binestLogPrior[handle, If[VectorQ[pts], {pts}, pts]]],
            "binestHandle" -> handle|>]
    ],
    "problemDef", Function[inferenceObject[$Failed]] (* BS:308 *)
];

(* ------------------------------------------------------------------ Gaussian processes (GP:201-422) *)
dataMatrix[v_?VectorQ] := List /@ N[v];
dataMatrix[m_?MatrixQ] := N[m];
defineGaussianProcess::outputDim = "Output data has has dimensions `1`. Only 1D output data is supported for GP regression at this time."; (* GP:208 *)
defineGaussianProcess[Rule[in_, out_], squaredExponentialKernel[sf_Symbol, ell_Symbol], nuggetVariance[sn_Symbol],
        None | 0 | 0., variables : {paramSpecPattern ..}, prior_, rest___Rule] := With[{
    x = dataMatrix[in], y = dataMatrix[out]},
    Which[
        Last[Dimensions[y]] =!= 1, Message[defineGaussianProcess::outputDim, Dimensions[y]]; inferenceObject[$Failed], (* GP:219-225 *)
        Length[x] =!= Length[y], inferenceObject[$Failed], (* GP:250-252 *)
        True, defineInferenceProblem[
            "Data" -> (x -> y), "PriorDistribution" -> prior, "Parameters" -> variables,
            "GeneratingDistribution" -> gpOperator[sf, ell, sn],
            "GaussianProcessData" -> <|"ModelFunctions" -> <|"KernelFunction" -> squaredExponentialKernel[sf, ell],
                "NuggetFunction" -> nuggetVariance[sn], "MeanFunction" -> (0 &)|>|>, (* GP:312-320 *)
            rest]]];

predictFromGaussianProcess[inferenceObject[result_?(AssociationQ[#] && KeyExistsQ[#, "GaussianProcessData"] && KeyExistsQ[#, "Samples"] && KeyExistsQ[#, "Data"] &)],
        n_Integer /; n > 1] := predictFromGaussianProcess[inferenceObject[result],
    CoordinateBoundsArray[CoordinateBounds[result["Data"][[1]]], Into[n - 1]]]; (* GP:332-342 *)
predictFromGaussianProcess[inferenceObject[result_?(AssociationQ[#] && KeyExistsQ[#, "GaussianProcessData"] && KeyExistsQ[#, "Samples"] &)],
        pts_List] := Module[{inputs, weights, ms},
    inputs = DeleteDuplicates @ With[{flat = If[ArrayDepth[pts] > 2, Flatten[pts, ArrayDepth[pts] - 2], pts]}, dataMatrix[flat]];
    weights = result["Samples", "CrudePosteriorWeight"]; (* GP:353; "Samples" is columnar in this host *)
    ms = binestGPPredict[result["binestHandle"], N @ result["Samples", "Point"], inputs];
    (
        AssociationThread[inputs,
            MapThread[Function[{mu, sd}, MixtureDistribution[weights, MapThread[NormalDistribution, {mu, sd}]]], (* GP:357, 401-418 *)
                {Transpose[ms[[1]]], Transpose[ms[[2]]]}]]
    ) /; ArrayQ[ms, 3, NumericQ]
];

(* ------------------------------------------------------------------ predictiveDistribution (BS:1373-1483) *)
predictiveDistribution::MissGenDist = "No generating distribution specified";
predictiveDistribution::unsampled = "Posterior has not been sampled yet";
predictiveDistribution[inferenceObject[result_?(AssociationQ[#] && MissingQ[#["Samples"]] &)], ___] := (
    Message[predictiveDistribution::unsampled]; $Failed);
predictiveDistribution[inferenceObject[result_?(AssociationQ[#] && !MissingQ[#["Samples"]] && MissingQ[#["GeneratingDistribution"]] &)], ___] := (
    Message[predictiveDistribution::MissGenDist]; $Failed);
(* point estimates keep one sample: "Samples" is columnar in this host, so take one position of every column *)
bestSample[result_, score_] := With[{i = First @ Ordering[score, -1]},
    Append[result, "Samples" -> Append[
        Map[If[AssociationQ[#], Map[Function[v, v[[{i}]]], #], #[[{i}]]] &, result["Samples"]],
        "CrudePosteriorWeight" -> {1.}]]];
predictiveDistribution[inferenceObject[result_?(AssociationQ[#] && !MissingQ[#["Samples"]] &)], rest___, "MaximumLikelihood"] :=
    predictiveDistribution[inferenceObject[bestSample[result, result["Samples", "LogLikelihood"]]], rest]; (* BS:1388-1402 *)
predictiveDistribution[inferenceObject[result_?(AssociationQ[#] && !MissingQ[#["Samples"]] &)], rest___, "MAP"] :=
    predictiveDistribution[inferenceObject[bestSample[result,
        result["Samples", "LogLikelihood"] + result["Samples", "LogPriorPDF"]]], rest]; (* BS:1404-1418 *)
predictiveDistribution[inferenceObject[result_?(AssociationQ[#] && ListQ[#["Data"]] && !MissingQ[#["Samples"]] && !MissingQ[#["GeneratingDistribution"]] &)]] :=
    With[{dist = Function[pt, result["GeneratingDistribution"] /. Thread[result["ParameterSymbols"] -> pt]]},
        MixtureDistribution[result["Samples", "CrudePosteriorWeight"], dist /@ result["Samples", "Point"]]]; (* BS:1420-1435 *)
predictiveDistribution[fst_, inputs_?VectorQ] := predictiveDistribution[fst, List /@ N[inputs], inputs]; (* BS:1437-1441 *)
predictiveDistribution[fst_, inputs_?MatrixQ] := predictiveDistribution[fst, N[inputs], inputs];         (* BS:1443-1446 *)
predictiveDistribution[inferenceObject[result_?(AssociationQ[#] && MatchQ[#["Data"], _Rule] && !MissingQ[#["Samples"]] && !MissingQ[#["GeneratingDistribution"]] &)],
        inputs_?MatrixQ, keys_List] /; Length[keys] === Length[inputs] := Module[{comp, w = result["Samples", "CrudePosteriorWeight"]},
    comp = binestPredictiveComponents[result["binestHandle"], N @ result["Samples", "Point"], inputs]; (* M x Q x C *)
    (
        AssociationThread[keys,
            Map[Function[tab, MixtureDistribution[w,
                    If[Last[Dimensions[comp]] === 2 && !MatchQ[result["GeneratingDistribution"], _categoricalSoftmax],
                        NormalDistribution @@@ tab,
                        Map[CategoricalDistribution[Range[Length[#]], #] &, tab]]]],
                Transpose[comp, {2, 1, 3}]]]
    ) /; ArrayQ[comp, 3, NumericQ]
]; (* BS:1448-1483 *)

(* ------------------------------------------------------------------ posterior sampler (BS:630-703) *)
createMCMCChain::start = "Please specify a starting point"; (* BS:650 *)
Options[createMCMCChain] = {"CovarianceLearnDelay" -> 20, "InitialCovariance" -> 1, "Seed" -> 1}; (* BS:698-701 *)
createMCMCChain[obj_?inferenceObjectQ, opts : OptionsPattern[]] /; !MatrixQ[obj["StartingPoints"], NumericQ] := (
    Message[createMCMCChain::start]; inferenceObject[$Failed]); (* BS:651-655 *)
createMCMCChain[obj_?inferenceObjectQ, opts : OptionsPattern[]] /; MatrixQ[obj["StartingPoints"], NumericQ] :=
    createMCMCChain[obj, First[obj["StartingPoints"]], opts]; (* BS:657-658 *)
createMCMCChain[obj_?inferenceObjectQ, startPt_?(VectorQ[#, NumericQ] &), opts : OptionsPattern[]] := With[{
    dim = Length[startPt]},
    With[{
        cov = Replace[OptionValue["InitialCovariance"], {   (* BS:676-683 *)
            n_?NumericQ :> DiagonalMatrix[ConstantArray[n, dim]],
            lst_?(VectorQ[#, NumericQ] &) /; Length[lst] === dim :> DiagonalMatrix[lst],
            Except[_?(MatrixQ[#, NumericQ] &)] :> DiagonalMatrix[ConstantArray[1, dim]]}],
        delay = Replace[OptionValue["CovarianceLearnDelay"], Except[_Integer] :> 20]}, (* BS:684-690 *)
        With[{h = check @ binestChainCreate[obj["binestHandle"], {N[startPt]}, N[cov], delay, OptionValue["Seed"]]},
            If[h === $Failed, inferenceObject[$Failed], markovChain[<|"Handle" -> h, "Dimension" -> dim|>]]]]];
markovChain[a_]["StateData"] := With[{v = binestChainState[a["Handle"], 1, a["Dimension"]], d = a["Dimension"]},
    {v[[;; d]], Round[v[[2 d + d^2 + 1]]], v[[d + 1 ;; 2 d]], Partition[v[[2 d + 1 ;; 2 d + d^2]], d]}]; (* {x, t, mean, cov} *)
markovChain[a_]["AcceptanceRate"] := With[{v = binestChainState[a["Handle"], 1, a["Dimension"]], d = a["Dimension"]},
    v[[2 d + d^2 + 2]]/Max[v[[2 d + d^2 + 1]] - 1, 1]];
iterateMCMC[markovChain[a_], n_Integer?Positive] := binestChainIterate[a["Handle"], n, 1, a["Dimension"]][[All, 1]]; (* BS:703 *)
iterateMCMC[markovChain[a_], {n_Integer?Positive, thin_Integer?Positive}] :=
    binestChainIterate[a["Handle"], n * thin, 1, a["Dimension"]][[thin ;; ;; thin, 1]];

(* ------------------------------------------------------------------ Laplace evidence (LaplaceApproximation.wl:22-30, 177-238) *)
laplaceLogEvidence[max_?NumericQ, prec_?(MatrixQ[#, NumericQ] &)] := With[{det = Det[prec]},
    max + (Length[prec] * Log[2 * Pi] - Log[det])/2 /; TrueQ[det > 0]];
laplaceLogEvidence[__] := Missing[];
approximateEvidence::nonposdef = "The Hessian at the maximum `1` is not negative definite."; (* LA:214-216 *)
approximateEvidence[inferenceObject[assoc_?AssociationQ]] := Module[{
    h = assoc["binestHandle"], names = assoc["ParameterSymbols"], d, logPost, start, max, mean, step, pts, v, hess, prec},
    d = Length[names];
    (* all stencil points of one evaluation go through ONE batched library call *)
    logPost[m_?(MatrixQ[#, NumericQ] &)] := binestLogLike[h, m] + binestLogPrior[h, m];
    logPost[p_?(VectorQ[#, NumericQ] &)] := First @ logPost[{p}];
    start = If[KeyExistsQ[assoc, "Samples"],
        assoc["Samples", "Point"][[First @ Ordering[assoc["Samples", "LogLikelihood"] + assoc["Samples", "LogPriorPDF"], -1]]],
        With[{c = binestSamplePrior[h, 4096, 77, 0]}, c[[First @ Ordering[logPost[c], -1]]]]];
    max = FindMaximum[logPost[Array[\[FormalX], d]], Transpose[{Array[\[FormalX], d], start}]]; (* LA:193-201 *)
    mean = Array[\[FormalX], d] /. Last[max];
    step = If[KeyExistsQ[assoc, "Samples"] && KeyExistsQ[assoc["Samples"], "CrudePosteriorWeight"],
        0.05 * Sqrt @ Diagonal @ Covariance @ WeightedData[assoc["Samples", "Point"], assoc["Samples", "CrudePosteriorWeight"]],
        10.^-4 * Abs[mean] + 10.^-6];
    pts = Flatten[Table[mean + si * step[[i]] * UnitVector[d, i] + sj * step[[j]] * UnitVector[d, j],
        {i, d}, {j, d}, {si, {1, -1}}, {sj, {1, -1}}], 3];
    v = ArrayReshape[logPost[pts], {d, d, 2, 2}];
    hess = Table[(v[[i, j, 1, 1]] - v[[i, j, 1, 2]] - v[[i, j, 2, 1]] + v[[i, j, 2, 2]])/(4 * step[[i]] * step[[j]]), {i, d}, {j, d}];
    prec = -(hess + Transpose[hess])/2; (* LA:209-213: minus the Hessian of the log posterior *)
    If[!PositiveDefiniteMatrixQ[prec], Message[approximateEvidence::nonposdef, mean]];
    <|"LogEvidence" -> laplaceLogEvidence[First[max], prec], "Maximum" -> {First[max], Thread[names -> mean]},
        "Mean" -> mean, "PrecisionMatrix" -> prec, "Parameters" -> names|> (* LA:219-234 *)
];

generateStartingPoints[inferenceObject[assoc_?AssociationQ], n_Integer, seed_Integer : 1] :=
    inferenceObject[Append[assoc, "StartingPoints" -> binestSamplePrior[assoc["binestHandle"], n, seed, 0]]]; (* BS:1046-1068 *)

(* ------------------------------------------------------------------ options: BS:833-851, 1366-1371 *)
Options[evidenceSampling] = {"PostProcessSamplingRuns" -> 100, "EmpiricalPosteriorDistributionType" -> "Simple", "Seed" -> 1};
Options[nestedSampling] = Join[{
    "SamplePoolSize" -> 100, "StartingPoints" -> Automatic, "MaxIterations" -> 10000, "MinIterations" -> 100,
    "MonteCarloMethod" -> Automatic, "MonteCarloSteps" -> 200, "TerminationFraction" -> 0.01, "Monitor" -> True,
    "LogLikelihoodMaximum" -> Automatic, "MinMaxAcceptanceRate" -> {0, 1},
    "BatchSize" -> 1 (* live points replaced per iteration; 1 = the reference scheme BS:980-1018 *)},
    Options[evidenceSampling]];
Options[parallelNestedSampling] = Join[DeleteCases[Options[nestedSampling], "StartingPoints" -> _], {"ParallelRuns" :> 4}];
Options[combineRuns] = Join[Options[evidenceSampling], {"MergeScheme" -> Automatic}];

createRunGroup[handle_, opts_, nRuns_, firstRun_, start_] := check @ binestRunCreate[handle,
    {opts["SamplePoolSize"], opts["BatchSize"], opts["MonteCarloSteps"], opts["MaxIterations"], opts["MinIterations"],
        opts["Seed"], firstRun, nRuns},
    N @ {opts["TerminationFraction"], opts["MinMaxAcceptanceRate"][[1]], opts["MinMaxAcceptanceRate"][[2]],
        (* "LogLikelihoodMaximum" -> number replaces the running maximum in the termination estimate (BS:925-932);
           Automatic travels as 1.*^308 (a packed real vector cannot carry NaN portably) *)
        If[NumericQ[opts["LogLikelihoodMaximum"]], opts["LogLikelihoodMaximum"], 1.*^308]},
    start];

runGroup[handle_, d_, opts_, nRuns_, firstRun_, start_] := Module[{run, fin, tables},
    run = createRunGroup[handle, opts, nRuns, firstRun, start];
    If[run === $Failed, Return[$Failed, Module]];
    fin = binestRunAdvance[run, 0];
    tables = Table[binestRunFetch[run, i, d], {i, 0, nRuns - 1}];
    binestRunFree[run];
    tables
];

(* fetch matrix -> association of the reference's result keys (BS:1026-1032) *)
resultAssociation[table_, d_, n_] := With[{rows = Most[table]},
    <|
        "Samples" -> <|
            "Point" -> rows[[All, ;; d]], "LogLikelihood" -> rows[[All, d + 1]], "LogPriorPDF" -> rows[[All, d + 2]],
            "AcceptanceRate" -> Replace[rows[[All, d + 3]], x_ /; !NumericQ[x] || x != x :> Missing["InitialSample"], {1}], (* BS:911 *)
            "PoolSize" -> Round @ rows[[All, d + 4]], "LogX" -> rows[[All, d + 5]], "X" -> Exp[rows[[All, d + 5]]],
            "CrudeLogPosteriorWeight" -> rows[[All, d + 6]]|>,
        "SamplePoolSize" -> n, "GeneratedNestedSamples" -> Length[rows] - n, "TotalSamples" -> Length[rows],
        "ParameterRanges" -> CoordinateBounds[rows[[All, ;; d]]]
    |>
];

nestedSampling[inferenceObject[assoc_?AssociationQ], opts : OptionsPattern[]] := Module[{
    o = Association[Options[nestedSampling], opts], start, d = Length[assoc["Parameters"]], tables},
    start = Replace[o["StartingPoints"], Automatic :> Lookup[assoc, "StartingPoints", {}]];
    If[MatrixQ[start, NumericQ], o["SamplePoolSize"] = Length[start]; start = {N[start]}, start = {}]; (* BS:1116-1131 *)
    tables = runGroup[assoc["binestHandle"], d, o, 1, 0, start];
    If[tables === $Failed, Return["Bad likelihood function", Module]]; (* BS:920 *)
    evidenceSampling[
        inferenceObject[Join[assoc, resultAssociation[First[tables], d, o["SamplePoolSize"]]]],
        Sequence @@ FilterRules[Normal[o], Options[evidenceSampling]]]
];

meanAndError[v_?VectorQ] := <|"Mean" -> Mean[v], "StandardError" -> StandardDeviation[v]|>; (* BS:1138-1149 *)

evidenceSampling[inferenceObject[assoc_?AssociationQ], opts : OptionsPattern[]] := Module[{
    o = Association[Options[evidenceSampling], opts],
    (* columns a previous evidenceSampling derived are recomputed: evidenceSampling[obj] re-post-processes a result, BS:1158-1160 *)
    s = KeyDrop[assoc["Samples"], {"SampledLogX", "LogPosteriorWeight", "CrudePosteriorWeight", "CrudeLogPosteriorWeight", "X", "LogX"}],
    n = Lookup[assoc, "binestLiveBlock", assoc["SamplePoolSize"]], (* combineRuns "PoolSizes": the tail that acts as the live set *)
    m, d, ord, cw, ev, r, pool, out},
    ord = Ordering[Transpose[{s["LogLikelihood"], s["Point"]}]]; (* SortBy {logL, point}, BS:814 *)
    s = Map[#[[ord]] &, s];
    m = Length[ord]; d = Length[First[s["Point"]]];
    pool = Lookup[s, "PoolSize", Join[ConstantArray[n, m - n], Range[n, 1, -1]]];
    cw = binestCrudeWeights[s["LogLikelihood"], pool, n];
    s["LogX"] = cw[[;; m]]; s["X"] = Exp[s["LogX"]]; s["CrudeLogPosteriorWeight"] = cw[[m + 1 ;; 2 m]];
    out = Join[assoc, <|"CrudeLogEvidence" -> cw[[2 m + 1]], "CrudeRelativeEntropy" -> cw[[2 m + 2]],
        "LogLikelihoodMaximum" -> cw[[2 m + 3]], "LogEstimatedMissingEvidence" -> cw[[2 m + 4]]|>]; (* BS:1183-1194 *)
    r = o["PostProcessSamplingRuns"];
    If[!TrueQ[IntegerQ[r] && r > 0], Return[inferenceObject[Append[out, "Samples" -> s]], Module]]; (* BS:1195-1197 *)
    ev = binestEvidenceSampling[s["Point"], s["LogLikelihood"], pool, n, r, o["Seed"]];
    s["CrudeLogPosteriorWeight"] -= out["CrudeLogEvidence"];                                   (* BS:1236 *)
    s["CrudePosteriorWeight"] = Exp[s["CrudeLogPosteriorWeight"]];                             (* BS:1237 *)
    With[{off = 2 r + r d},
        s["LogPosteriorWeight"] = <|"Mean" -> ev[[off + 1 ;; off + m]], "StandardError" -> ev[[off + m + 1 ;; off + 2 m]]|>;
        s["SampledLogX"] = <|"Mean" -> ev[[off + 2 m + 1 ;; off + 3 m]], "StandardError" -> ev[[off + 3 m + 1 ;; off + 4 m]]|>];
    ord = Ordering[-s["CrudeLogPosteriorWeight"]];                                            (* BS:1241 *)
    s = Map[If[AssociationQ[#], Map[Function[v, v[[ord]]], #], #[[ord]]] &, s];
    inferenceObject @ Join[out, <|
        "Samples" -> s,
        "LogEvidence" -> meanAndError[ev[[;; r]]],                                            (* BS:1254 *)
        "RelativeEntropy" -> meanAndError[ev[[r + 1 ;; 2 r]]],                                 (* BS:1263 *)
        "ParameterExpectedValues" -> AssociationThread[
            Lookup[assoc, "ParameterSymbols", Range[d]],
            meanAndError /@ Transpose[Partition[ev[[2 r + 1 ;; 2 r + r d]], d]]],               (* BS:1255-1262 *)
        "EmpiricalPosteriorDistribution" -> EmpiricalDistribution[s["CrudePosteriorWeight"] -> s["Point"]] (* BS:1272-1277 *)
    |>]
];

(* samples of one result sorted by {logL, point} (BS:814): {levels, pool sizes} *)
sortedLevelsAndPools[a_Association] := With[{s = a["Samples"], n = a["SamplePoolSize"]},
    With[{ord = Ordering[Transpose[{s["LogLikelihood"], s["Point"]}]], k = Length[s["LogLikelihood"]]},
        {s["LogLikelihood"][[ord]],
         If[KeyExistsQ[s, "PoolSize"], Round[s["PoolSize"][[ord]]], Join[ConstantArray[n, Max[k - n, 0]], Range[Min[n, k], 1, -1]]]}]];

(* the reference's pool structure: n for every deleted point, then the live set n..1 (BS:785-799) *)
referencePoolQ[a_Association] := With[{lp = sortedLevelsAndPools[a], n = a["SamplePoolSize"]},
    With[{p = Last[lp], k = Length[Last[lp]]},
        k >= n && p[[;; k - n]] === ConstantArray[n, k - n] && p[[k - n + 1 ;;]] === Range[n, 1, -1]]];

(* BS:1293-1315.  "MergeScheme" (not in the reference): "Reference" = the literal formula, pool size
   Total[SamplePoolSize] for the first M - nTot samples and the nTot best as a live set (BS:1307-1309 -> BS:785-799);
   "PoolSizes" = at every sample the SUM over runs of that run's pool size at the sample's likelihood level, used
   consistently to the last sample (needed when runs replaced "BatchSize" > 1 points per iteration);
   Automatic = "Reference" when every run has the reference's pool structure, else "PoolSizes". *)
combineRuns[results : inferenceObject[_?AssociationQ] .., opts : OptionsPattern[]] := Module[{
    assocs = {results}[[All, 1]], cols = {"Point", "LogLikelihood", "LogPriorPDF", "AcceptanceRate"},
    scheme = OptionValue["MergeScheme"], pools, nTot, joined, keep, ord, merged, m, lps, base, levels, deltas, eo, sl, csum,
    rank, below, pool, dis, live, extra = <||>},
    pools = #["SamplePoolSize"] & /@ assocs; nTot = Total[pools];
    joined = AssociationMap[Function[c, Join @@ (#["Samples"][c] & /@ assocs)], cols];
    keep = Sort[First /@ Values[PositionIndex[joined["Point"]]]];                              (* DeleteDuplicatesBy Point: first kept, BS:1294-1297 *)
    joined = Map[#[[keep]] &, joined];
    ord = Ordering[Transpose[{joined["LogLikelihood"], joined["Point"]}]];                     (* SortBy {logL, point} *)
    merged = Map[#[[ord]] &, joined];
    m = Length[ord];
    If[scheme === Automatic, scheme = If[AllTrue[assocs, referencePoolQ], "Reference", "PoolSizes"]];
    If[scheme === "Reference",
        merged["PoolSize"] = Join[ConstantArray[nTot, Max[m - nTot, 0]], Range[Min[nTot, m], 1, -1]],
        (* every run's pool size is a step function of the likelihood level that changes at its own samples: the sum
           over runs is one cumulative sum over all samples sorted by level *)
        lps = sortedLevelsAndPools /@ assocs;
        base = Total[First[Last[#]] & /@ lps];
        levels = Join @@ (First /@ lps);
        deltas = Join @@ (Differences[Append[Last[#], 0]] & /@ lps);
        eo = Ordering[levels]; sl = levels[[eo]]; csum = Prepend[Accumulate[deltas[[eo]]], 0];
        (* below[[i]] = number of run samples strictly below merged level i: joint ranking, merged entries first on ties *)
        rank = Ordering[Ordering[Join[Transpose[{merged["LogLikelihood"], ConstantArray[0, m]}], Transpose[{sl, ConstantArray[1, Length[sl]]}]]]];
        below = rank[[;; m]] - Range[m];
        pool = base + csum[[below + 1]];
        merged["PoolSize"] = pool;
        (* once every run is inside its final live set the sum falls by one per sample: that tail is the live set of
           the merged run in the sense of BS:791-797 *)
        dis = Flatten @ Position[Unitize[pool - Range[m, 1, -1]], 1];
        live = If[dis === {}, m, m - Last[dis]];
        extra = <|"binestLiveBlock" -> Max[live, 1]|>
    ];
    evidenceSampling[
        inferenceObject @ Join[KeyDrop[First[assocs], "binestLiveBlock"], <|
            "Samples" -> merged,
            "LogLikelihoodMaximum" -> Max[#["LogLikelihoodMaximum"] & /@ assocs],             (* BS:1306 *)
            "SamplePoolSize" -> nTot, "GeneratedNestedSamples" -> m - nTot, "TotalSamples" -> m, (* BS:1307-1309 *)
            "MergeScheme" -> scheme|>, extra],
        Sequence @@ FilterRules[{opts}, Options[evidenceSampling]]]
];

parallelNestedSampling::startingPts = "Cannot use pre-specified starting points for parallel sampling because each parallel process should generate starting points independently.
Continuing with option \"SamplePoolSize\" -> `1`"; (* BS:1317-1318 *)

parallelNestedSampling[inferenceObject[assoc_?AssociationQ], opts : OptionsPattern[]] /; MatrixQ[assoc["StartingPoints"], NumericQ] := (
    Message[parallelNestedSampling::startingPts, Length[assoc["StartingPoints"]]];
    parallelNestedSampling[inferenceObject[KeyDrop[assoc, "StartingPoints"]], "SamplePoolSize" -> Length[assoc["StartingPoints"]], opts]
); (* BS:1320-1332 *)

(* One library call advances all runs in lock step on the GPU (instead of ParallelTable over subkernels, BS:1349-1357:
   eight subkernels would mean eight CUDA contexts fighting for the device). *)
parallelNestedSampling[inferenceObject[assoc_?AssociationQ], opts : OptionsPattern[]] := Module[{
    o = Association[Options[parallelNestedSampling], opts], d = Length[assoc["Parameters"]], tables, run, fin, r, tab, m, nTot,
    rows, draws, last, s, scheme, run0},
    r = o["PostProcessSamplingRuns"];
    If[!TrueQ[IntegerQ[r] && r > 0],
        (* no Monte-Carlo post-processing asked for (BS:1195-1197): per-run fetch and the host merge *)
        tables = runGroup[assoc["binestHandle"], d, o, o["ParallelRuns"], 0, {}];
        If[tables === $Failed, Return["Bad likelihood function", Module]];
        Return[combineRuns[
            Sequence @@ (inferenceObject[Join[assoc, resultAssociation[#, d, o["SamplePoolSize"]]]] & /@ tables),
            Sequence @@ FilterRules[Normal[o], Options[combineRuns]]], Module]];
    (* combineRuns -> evidenceSampling of the group's runs in ONE library call on the device (binest_run_combine,
       csrc/merge.cu): same result as combineRuns[runs...] above, the runs never travel to the host one by one *)
    run = createRunGroup[assoc["binestHandle"], o, o["ParallelRuns"], 0, {}];
    If[run === $Failed, Return["Bad likelihood function", Module]];
    fin = binestRunAdvance[run, 0];
    scheme = If[o["BatchSize"] == 1, "Reference", "PoolSizes"]; (* combineRuns' Automatic: K = 1 runs have the reference's pool structure *)
    tab = binestRunCombine[run, d, If[scheme === "Reference", 0, 1], Max[r, 2], o["Seed"]];
    binestRunFree[run];
    If[!MatrixQ[tab], Return[$Failed, Module]];
    last = Last[tab]; m = Round[last[[6]]]; nTot = o["ParallelRuns"] o["SamplePoolSize"];
    rows = tab[[;; m]]; draws = tab[[m + 1 ;; m + Max[r, 2]]];
    s = <|
        "Point" -> rows[[All, ;; d]], "LogLikelihood" -> rows[[All, d + 1]], "LogPriorPDF" -> rows[[All, d + 2]],
        "AcceptanceRate" -> Replace[rows[[All, d + 3]], x_ /; !NumericQ[x] || x != x :> Missing["InitialSample"], {1}], (* BS:911 *)
        "LogX" -> rows[[All, d + 4]], "X" -> rows[[All, d + 5]],
        "CrudeLogPosteriorWeight" -> rows[[All, d + 6]], "CrudePosteriorWeight" -> rows[[All, d + 7]],
        "SampledLogX" -> <|"Mean" -> rows[[All, d + 8]], "StandardError" -> rows[[All, d + 9]]|>,
        "LogPosteriorWeight" -> <|"Mean" -> rows[[All, d + 10]], "StandardError" -> rows[[All, d + 11]]|>,
        "PoolSize" -> Round @ rows[[All, d + 12]], "RunIndex" -> Round @ rows[[All, d + 13]]|>;
    run0 = Pick[s["Point"], s["RunIndex"], 0]; (* combineRuns keeps First[results] for the other keys (BS:1299-1305) *)
    inferenceObject @ Join[assoc, <|
        "Samples" -> s,
        "ParameterRanges" -> CoordinateBounds[run0],
        "SamplePoolSize" -> nTot, "GeneratedNestedSamples" -> m - nTot, "TotalSamples" -> m, "MergeScheme" -> scheme, (* BS:1307-1309 *)
        "CrudeLogEvidence" -> last[[1]], "CrudeRelativeEntropy" -> last[[2]], "LogLikelihoodMaximum" -> last[[3]],
        "LogEstimatedMissingEvidence" -> last[[4]],                                                  (* BS:1183-1194 *)
        "LogEvidence" -> meanAndError[draws[[All, 1]]],                                             (* BS:1254 *)
        "RelativeEntropy" -> meanAndError[draws[[All, 2]]],                                          (* BS:1263 *)
        "ParameterExpectedValues" -> AssociationThread[
            Lookup[assoc, "ParameterSymbols", Range[d]], meanAndError /@ Transpose[draws[[All, 3 ;; d + 2]]]], (* BS:1255-1262 *)
        "EmpiricalPosteriorDistribution" -> EmpiricalDistribution[s["CrudePosteriorWeight"] -> s["Point"]]|>, (* BS:1272-1277 *)
        If[scheme === "Reference", <||>, <|"binestLiveBlock" -> Round[last[[5]]]|>]]
];

End[];
EndPackage[];
