% make_reference_goldens.m - run the REFERENCE's own functions (COMPiLELab/krylov_robustness, functions/*.m) on the
% committed inputs tests/golden/reference_inputs.mat and write tests/golden/reference_golden.json.
%
% This image has neither MATLAB nor GNU Octave, so parity is "unpinned" (DESIGN.md section 2): nothing ties the NumPy/SciPy
% oracle to outputs of the reference itself.  Anyone with Octave (>= 6) or MATLAB and a checkout of the reference can
% lift that:
%     cd <this repo>;  octave --eval "refdir='/path/to/krylov_robustness'; run('scripts/make_reference_goldens.m')"
% then `pytest tests/test_reference_goldens.py` checks the oracle (CPU tier) and the device path (GPU tier) against the
% file at 1e-10 with equal iteration counts.  Only built-ins that Octave has are used (no table/graph/funm: the cases
% below use @exp, for which the reference calls expm).  mc_trace draws its probes from randn (mc_trace.m:43-44,
% unseeded); a shim randn.m placed first on the path replays the committed probe blocks instead.
% Drop-in mode: with `dropin_dir` set (to krylov_robustness_b200/matlab) the SAME calls go through this repository's
% wrappers -> MEX gateway -> CUDA instead, and the sections that need reference files without a wrapper (greedy_krylov,
% find_top_*, edge2low_rank, compute_centrality, multiple_frechet_eval: host glue above the path) are skipped;
% tests/test_gpu_dropin_matlab.py compares what comes out with the reference's numbers.
dropin = exist('dropin_dir', 'var');
if ~dropin && ~exist('refdir', 'var'), error('set refdir to a checkout of COMPiLELab/krylov_robustness'); end
here = fileparts(mfilename('fullpath'));
root = fileparts(here);
in = load(fullfile(root, 'tests', 'golden', 'reference_inputs.mat'));
if dropin, addpath(dropin_dir); else, addpath(fullfile(refdir, 'functions')); end
out = struct();

% ---- trace_fun_update on candidate edges (functions/krylov_miobi.m:76-99 call shape)
n0 = size(in.A0, 1);
for kind = {'break', 'make'}
    E = in.(['A0_' kind{1} '_edges']);
    sgn = -1; if strcmp(kind{1}, 'make'), sgn = 1; end
    x = zeros(size(E, 1), 1); it = x; lk = x;
    for h = 1:size(E, 1)
        U = zeros(n0, 2); U(E(h, 1), 1) = 1; U(E(h, 2), 2) = 1;
        [x(h), it(h), lk(h)] = trace_fun_update(in.A0, U, sgn * [0 1; 1 0], in.A0_tol, 100, 0);
    end
    out.(['A0_' kind{1} '_x']) = x; out.(['A0_' kind{1} '_iter']) = it; out.(['A0_' kind{1} '_lucky']) = lk;
end
nR = size(in.Rome, 1);
x = zeros(size(in.Rome_edges, 1), 1); it = x;
for h = 1:size(in.Rome_edges, 1)
    U = zeros(nR, 2); U(in.Rome_edges(h, 1), 1) = 1; U(in.Rome_edges(h, 2), 2) = 1;
    [x(h), it(h)] = trace_fun_update(in.Rome, U, -[0 1; 1 0], in.Rome_tol, 100, 0, @sinh);
end
out.Rome_sinh_x = x; out.Rome_sinh_iter = it;

% ---- fun_update, both variants (nargout decides, functions/fun_update.m:69,77)
% With three outputs the reference takes the Lanczos variant, keeps only the two-block window in Um and then
% executes `Um = Um(:, 1:size(Xm, 1))` (fun_update.m:137), which is out of range from the third step on: the
% three-output form of the reference ends in an index error.  Recorded as a fact of the reference.
out.Mexico_lanczos_form_raises = 0;
try
    [Xm, it3, lk3] = fun_update(in.Mexico, in.Mexico_U, in.Mexico_B, @exp, in.Mexico_tol, 100, 0);
    out.Mexico_lanczos_trace = trace(Xm); out.Mexico_lanczos_iter = it3; out.Mexico_lanczos_dim = size(Xm, 1);
catch err
    out.Mexico_lanczos_form_raises = 1;
end
[Xm, it4, lk4, Um] = fun_update(in.Mexico, in.Mexico_U, in.Mexico_B, @exp, in.Mexico_tol, 100, 0);
F = Um * Xm * Um';
out.Mexico_arnoldi_iter = it4; out.Mexico_arnoldi_dim = size(Xm, 1);
out.Mexico_arnoldi_update_diag = full(diag(F));       % basis independent: diag of f(A+UBU') - f(A)

% ---- function_multiple_entries, expmv, select_taylor_degree / normAm
[X, itE] = function_multiple_entries(in.A0, in.A0_omega, @exp, 1e-10 * exp(normest(in.A0)), 100, inf, 0);
out.A0_entries = X; out.A0_entries_iter = itE;
[f, s, m, mv, mvd, unA] = expmv(1, in.A0, in.A0_b, [], 'double');
out.A0_expmv_f = f(:); out.A0_expmv_info = [s m mv mvd unA];
[c9, mv9] = normAm(in.A0, 9);
out.A0_normAm9 = [c9 mv9];
out.A0_normest = normest(in.A0, 1e-2);

% ---- lanczos_krylov: Ritz values of the projection after 4 steps (basis-independent invariants)
[V, H, p] = lanczos_krylov(in.A0, in.A0_b);
for j = 2:4, [V, H, p] = lanczos_krylov(V, H, p); end
G = H(1:end - 3, :); out.A0_lanczos_ritz = sort(eig((G + G') / 2));

% ---- the operator plug-in point of the L1 functions (lanczos_krylov.m:32,78-79: A may be a struct with a `multiply`
%      field).  Drop-in mode only: kr_operator(A) builds that struct around the device SpMM; the keys it writes are
%      compared with A0_lanczos_ritz above and with A0 * b.
if dropin
    op = kr_operator(in.A0);
    [V, H, p] = lanczos_krylov(op, in.A0_b);
    for j = 2:4, [V, H, p] = lanczos_krylov(V, H, p); end
    G = H(1:end - 3, :); out.A0_lanczos_ritz_operator_struct = sort(eig((G + G') / 2));
    [V, K, H, p] = arnoldi_krylov(op, in.A0_b);
    G = H(1:end - 3, :); out.A0_arnoldi_first_block_operator_struct = sort(eig((G + G') / 2));
    y = op.multiply(1.0, 0.0, in.A0_b); out.A0_operator_multiply = y(:);
end

% ---- mc_trace with replayed probes
shim = tempname(); mkdir(shim);
fid = fopen(fullfile(shim, 'randn.m'), 'w');
fprintf(fid, 'function r = randn(varargin)\nglobal KR_PROBES KR_PROBE_POS\nr = KR_PROBES(:, KR_PROBE_POS + (1:10)); KR_PROBE_POS = KR_PROBE_POS + 10;\nend\n');
fclose(fid);
global KR_PROBES KR_PROBE_POS
KR_PROBES = in.A0_probes; KR_PROBE_POS = 0;
addpath(shim);
[tr, res, itm] = mc_trace(in.A0, n0, 1e-3, 60, 1, 0);
rmpath(shim);
out.A0_mc_trace = [tr res itm];

% ---- greedy_krylov: 5 rounds, Q = 50 (Tests/test_unweighted_break.m:74 call shape, smaller)
if ~dropin
    [edges, rob] = greedy_krylov(in.A0, 5, 50, in.A0_centrality, 'min', in.A0_tol, 100, inf, 0, 'break');
    out.A0_greedy_edges = edges(:); out.A0_greedy_rob = rob;
end

% ==== second batch =========================================================================================
% ---- trace_fun_update: self loop (rank one, krylov_miobi.m:88-99), dense branch n <= 130 (trace_fun_update.m:37-51),
%      an edge SET with rk > 2 in the call shape of Tests/test_unweighted_break.m:92-95 (full(U)), cosh
U = zeros(n0, 1); U(in.A0_self_node) = 1;
[xs, its, lks] = trace_fun_update(in.A0, U, -1, in.A0_tol, 100, 0);
out.A0_selfloop = [xs its lks];
As = in.A0(1:100, 1:100);
U = zeros(100, 2); U(3, 1) = 1; U(7, 2) = 1;
[xd, itd, lkd] = trace_fun_update(As, U, [0 1; 1 0], 1e-8, 100, 0);
out.A0_dense_branch = [xd itd lkd];
if ~dropin
    [Us, Bs] = edge2low_rank(in.A0_set_edges, n0);
else    % the same U, B written out (edge2low_rank.m:3-12) - the reference file is not on the drop-in path
    ut = unique(in.A0_set_edges(:)); Us = sparse(ut, 1:length(ut), 1, n0, length(ut)); Bs = zeros(length(ut));
    for q = 1:size(in.A0_set_edges, 1)
        a1 = find(ut == in.A0_set_edges(q, 1)); a2 = find(ut == in.A0_set_edges(q, 2)); Bs(a1, a2) = -1; Bs(a2, a1) = -1;
    end
end
[xe, ite, lke] = trace_fun_update(in.A0, full(Us), Bs, in.A0_tol, 100, 0);
out.A0_edge_set = [xe ite lke]; out.A0_edge_set_rk = size(Us, 2);
x = zeros(size(in.Rome_edges, 1), 1); it = x;
for h = 1:size(in.Rome_edges, 1)
    U = zeros(nR, 2); U(in.Rome_edges(h, 1), 1) = 1; U(in.Rome_edges(h, 2), 2) = 1;
    [x(h), it(h)] = trace_fun_update(in.Rome, U, -[0 1; 1 0], in.Rome_tol, 100, 0, @cosh);
end
out.Rome_cosh_x = x; out.Rome_cosh_iter = it;

% ---- fun_update with four outputs for sinh (fun_update.m:55-56)
[Xm, it5, lk5, Um] = fun_update(in.Mexico, in.Mexico_U, in.Mexico_B, @sinh, in.Mexico_tol, 100, 0);
F = Um * Xm * Um';
out.Mexico_arnoldi_sinh_iter = it5; out.Mexico_arnoldi_sinh_update_diag = full(diag(F));

% ---- arnoldi_krylov: Ritz values after 4 steps, orthogonality of the basis
[V, K, H, p] = arnoldi_krylov(in.A0, in.A0_b);
for j = 2:4, [V, K, H, p] = arnoldi_krylov(V, K, H, p); end
G = H(1:end - 3, :); out.A0_arnoldi_ritz = sort(eig((G + G') / 2));
out.A0_arnoldi_dims = [size(V) size(H) size(K)];

% ---- select_taylor_degree on its own; expmv with full_term and t ~= 1; the normest1 branch of normAm (a matrix with
%      self loops gets negative diagonal entries from the shift, normAm.m:16,24-26 and the nested afun_power)
[M, mvs, alpha, unA] = select_taylor_degree(in.A0, in.A0_b, [], [], 'double', true, false);
out.A0_std_M = M(:); out.A0_std_info = [mvs unA]; out.A0_std_alpha = alpha;
[f, s, m, mv, mvd, unA] = expmv(0.5, in.A0, in.A0_b, [], 'double', true, false, true);
out.A0_expmv_half_full_term_f = f(:); out.A0_expmv_half_full_term_info = [s m mv mvd unA];
Al = in.A0 + sparse(in.A0_loop_nodes, in.A0_loop_nodes, 1, n0, n0);
[f, s, m, mv, mvd, unA] = expmv(1, Al, in.A0_b, [], 'double');
out.A0_loops_expmv_f = f(:); out.A0_loops_expmv_info = [s m mv mvd unA];
mu = full(trace(Al)) / n0;
[c5, mv5] = normAm(Al - mu * speye(n0), 5);
out.A0_loops_normAm5 = [c5 mv5];

% ---- trace_exp (mc_trace around expmv, tol 1e-4, maxit 1000) with replayed probes
KR_PROBES = in.A0_probes_long; KR_PROBE_POS = 0;
addpath(shim);
tre = trace_exp(in.A0);
rmpath(shim);
out.A0_trace_exp = [tre KR_PROBE_POS];

if ~dropin
% ---- candidate generators, both orderings
E1 = find_top_edges(in.A0, in.A0_centrality, 30, 'min');  out.A0_top_edges_min = E1(:);
E2 = find_top_edges(in.A0, in.A0_centrality, 30, 'mult'); out.A0_top_edges_mult = E2(:);
E3 = find_top_missing_edges(in.A0, in.A0_centrality, 30, 'min');  out.A0_top_missing_min = E3(:);
E4 = find_top_missing_edges(in.A0, in.A0_centrality, 30, 'mult'); out.A0_top_missing_mult = E4(:);

% ---- greedy 'make'; krylov_miobi on an explicit list with a self loop and rescale ~= 1 (krylov_miobi.m:78-99)
[edges, rob] = greedy_krylov(in.A0, 3, 30, in.A0_centrality, 'min', in.A0_tol, 100, inf, 0, 'make');
out.A0_greedy_make_edges = edges(:); out.A0_greedy_make_rob = rob;
end
[edges, rob, Anew] = krylov_miobi(in.A0, 2, in.A0_mixed_edges, in.A0_tol, 100, inf, 0, 'break', 2);
out.A0_miobi_rescale_edges = edges(:); out.A0_miobi_rescale_rob = rob; out.A0_miobi_rescale_nnz = nnz(Anew);

% ---- weighted experiments: objective + gradient callbacks and exact Hessians (Tests/test_weighted_*_hessian.m)
tolE = 1e-10 * exp(normest(in.Mexico));
eA = function_multiple_entries(in.Mexico, in.Mexico_Omega, @exp, tolE, 100, inf, 0);
[fv, gr] = fun_and_grad_krylov_exp(in.Mexico_X, in.Mexico, in.Mexico_Omega, eA, 1e-10, 100, 0);
out.Mexico_fg_exp = [fv; gr]; out.Mexico_eA = eA;
[fv0, gr0] = fun_and_grad_krylov_exp(0 * in.Mexico_X, in.Mexico, in.Mexico_Omega, eA, 1e-10, 100, 0);
out.Mexico_fg_exp_at_zero = [fv0; gr0];
dfA = function_multiple_entries(in.Mexico, in.Mexico_Omega, @cosh, 1e-10 * cosh(normest(in.Mexico)), 100, inf, 0);
[fv, gr] = fun_and_grad_krylov_fun(in.Mexico_X, in.Mexico, in.Mexico_Omega, @sinh, @cosh, dfA, 1e-10, 100, 0);
out.Mexico_fg_sinh = [fv; gr]; out.Mexico_dfA_cosh = dfA;
Hes = hessianfcn_exp(in.Mexico_X, in.Mexico, in.Mexico_Omega, 1e-10, 100);
out.Mexico_hessian_exp = Hes(:);
Hes = hessianfcn_fun(in.Mexico_X, in.Mexico, in.Mexico_Omega, @sinh, 1e-10, 100);
out.Mexico_hessian_sinh = Hes(:);
% 30 modifiable edges (Tests/test_weighted_sinh_lbfgs.m call shape): wide blocks, dense fall-back of fun_update.m:84-90
dfA30 = function_multiple_entries(in.Mexico, in.Mexico_Omega30, @cosh, 1e-10 * cosh(normest(in.Mexico)), 100, inf, 0);
[fv, gr] = fun_and_grad_krylov_fun(in.Mexico_X30, in.Mexico, in.Mexico_Omega30, @sinh, @cosh, dfA30, 1e-8, 100, 0);
out.Mexico_fg30_sinh = [fv; gr]; out.Mexico_dfA30_cosh = dfA30;
eA30 = function_multiple_entries(in.Mexico, in.Mexico_Omega30, @exp, tolE, 100, inf, 0);
[fv, gr] = fun_and_grad_krylov_exp(in.Mexico_X30, in.Mexico, in.Mexico_Omega30, eA30, 1e-8, 100, 0);
out.Mexico_fg30_exp = [fv; gr]; out.Mexico_eA30 = eA30;
% BASELINE config C2 at the reference's literal call: Omega = ALL edges of the Anaheim road network.  U becomes an
% n x n selector, fun_update.m:84-90 takes its dense branch and trace_fun_update stops on a lucky breakdown after one
% block step.  (The device evaluates the same gradient sparsely, fun_and_grad_all_edges; the dense cosh(A) entries that
% the callback wants as dfA are formed here the way Tests/test_weighted_sinh_lbfgs.m does, from the dense function.)
fA = full(in.Anaheim);
Cd = (expm(fA) + expm(-fA)) / 2;
nA = size(fA, 1);
dfAn = Cd((in.Anaheim_Omega(:, 2) - 1) * nA + in.Anaheim_Omega(:, 1));
[fv, gr] = fun_and_grad_krylov_fun(in.Anaheim_X, in.Anaheim, in.Anaheim_Omega, @sinh, @cosh, dfAn, 1e-8, 100, 0);
out.Anaheim_all_edges_fg_sinh = [fv; gr];
if ~dropin
    [Umf, Xmf, Vmf, rowf, colf, itf] = multiple_frechet_eval(in.Mexico, in.Mexico_Omega, @exp, 1e-10, 100, inf, 0);
    out.Mexico_frechet_iter = itf;
end

if ~dropin
% ---- compute_centrality (eigs is ARPACK here as in MATLAB; the vector's sign is arbitrary, the reference takes abs)
out.A0_centrality_eig = compute_centrality(in.A0, 'eig');
out.A0_centrality_deg = full(compute_centrality(in.A0, 'deg'));
out.A0_centrality_pr = compute_centrality(in.A0, 'pr');
out.A0_centrality_exp = full(compute_centrality(in.A0, 'exp'));
end

% ---- JSON by hand (jsonencode is missing from older Octave)
if ~exist('golden_path', 'var'), golden_path = fullfile(root, 'tests', 'golden', 'reference_golden.json'); end
fid = fopen(golden_path, 'w');
names = fieldnames(out);
fprintf(fid, '{\n');
for k = 1:numel(names)
    v = double(out.(names{k})(:));
    fprintf(fid, '  "%s": [', names{k});
    for q = 1:numel(v)
        if q > 1, fprintf(fid, ', '); end
        fprintf(fid, '%.17g', v(q));
    end
    if k < numel(names), fprintf(fid, '],\n'); else, fprintf(fid, ']\n'); end
end
fprintf(fid, '}\n');
fclose(fid);
fprintf('wrote tests/golden/reference_golden.json (%d entries)\n', numel(names));
