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
if ~exist('refdir', 'var'), error('set refdir to a checkout of COMPiLELab/krylov_robustness'); end
here = fileparts(mfilename('fullpath'));
root = fileparts(here);
in = load(fullfile(root, 'tests', 'golden', 'reference_inputs.mat'));
addpath(fullfile(refdir, 'functions'));
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
[Xm, it3, lk3] = fun_update(in.Mexico, in.Mexico_U, in.Mexico_B, @exp, in.Mexico_tol, 100, 0);
out.Mexico_lanczos_trace = trace(Xm); out.Mexico_lanczos_iter = it3; out.Mexico_lanczos_dim = size(Xm, 1);
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
[edges, rob] = greedy_krylov(in.A0, 5, 50, in.A0_centrality, 'min', in.A0_tol, 100, inf, 0, 'break');
out.A0_greedy_edges = edges(:); out.A0_greedy_rob = rob;

% ---- JSON by hand (jsonencode is missing from older Octave)
fid = fopen(fullfile(root, 'tests', 'golden', 'reference_golden.json'), 'w');
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
