% make_reference_goldens_c1.m - BASELINE config C1 at the reference's own call shape (Tests/test_unweighted_break.m:74,
% Tests/test_unweighted_make.m): the whole greedy run, k = 50 rounds of Q = 250 candidates on Oregon A0 ('break'), ten
% 'make' rounds, and three rounds on the largest Oregon graph (A7), executed by the REFERENCE's functions/*.m.
% Same conventions as make_reference_goldens.m (set refdir; run under Octave / MATLAB or through
% `python scripts/run_reference_goldens.py --c1`).  12 500 + 2 500 + 750 calls of trace_fun_update: ~10 minutes under
% the interpreter of oracle/mlab, so the result is committed (tests/golden/reference_golden_c1.json) and not re-derived
% by the test tier.  Which of two structurally equivalent candidates (two leaves of one hub) wins a round is decided in
% the last bits of the scores: this file records what the reference's arithmetic picks.
if ~exist('refdir', 'var'), error('set refdir to a checkout of COMPiLELab/krylov_robustness'); end
here = fileparts(mfilename('fullpath'));
root = fileparts(here);
in = load(fullfile(root, 'tests', 'golden', 'reference_inputs.mat'));
in1 = load(fullfile(root, 'tests', 'golden', 'reference_inputs_c1.mat'));
addpath(fullfile(refdir, 'functions'));
out = struct();
[edges, rob] = greedy_krylov(in.A0, 50, 250, in.A0_centrality, 'min', in.A0_tol, 100, inf, 0, 'break');
out.C1_A0_break_k50_Q250_edges = edges(:); out.C1_A0_break_k50_Q250_rob = rob;
[edges, rob] = greedy_krylov(in.A0, 10, 250, in.A0_centrality, 'min', in.A0_tol, 100, inf, 0, 'make');
out.C1_A0_make_k10_Q250_edges = edges(:); out.C1_A0_make_k10_Q250_rob = rob;
[edges, rob] = greedy_krylov(in1.A7, 3, 250, in1.A7_centrality, 'min', in1.A7_tol, 100, inf, 0, 'break');
out.C1_A7_break_k3_Q250_edges = edges(:); out.C1_A7_break_k3_Q250_rob = rob;
if ~exist('golden_path', 'var'), golden_path = fullfile(root, 'tests', 'golden', 'reference_golden_c1.json'); end
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
