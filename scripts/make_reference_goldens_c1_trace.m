% make_reference_goldens_c1_trace.m - the other half of BASELINE config C1: trace_exp (trace_exp.m -> mc_trace.m ->
% expmv.m, tol 1e-4, maxit 1000) in parity mode on Oregon A0, A4, A7 (the largest) and A8, executed by the REFERENCE's
% functions/*.m.  mc_trace draws its probes from randn (mc_trace.m:43-44, unseeded); here a shim randn.m replays a
% block of +-1 signs from one Park-Miller generator per column, which MATLAB doubles and NumPy integers reproduce bit for
% bit (scripts/make_reference_inputs_c1.py::lcg_sign_probes), so nothing large has to be stored.
% Same conventions as make_reference_goldens.m; `python scripts/run_reference_goldens.py --which c1trace`.
if ~exist('refdir', 'var'), error('set refdir to a checkout of COMPiLELab/krylov_robustness'); end
here = fileparts(mfilename('fullpath'));
root = fileparts(here);
in = load(fullfile(root, 'tests', 'golden', 'reference_inputs.mat'));
in1 = load(fullfile(root, 'tests', 'golden', 'reference_inputs_c1.mat'));
addpath(fullfile(refdir, 'functions'));
shim = tempname(); mkdir(shim);
fid = fopen(fullfile(shim, 'randn.m'), 'w');
fprintf(fid, 'function r = randn(varargin)\nglobal KR_PROBES KR_PROBE_POS\nr = KR_PROBES(:, KR_PROBE_POS + (1:10)); KR_PROBE_POS = KR_PROBE_POS + 10;\nend\n');
fclose(fid);
global KR_PROBES KR_PROBE_POS
out = struct();
graphs = {in.A0, in1.A4, in1.A7, in1.A8};
names = {'A0', 'A4', 'A7', 'A8'};
for g = 1:4
    A = graphs{g};
    n = size(A, 1);
    x = 1:680;
    for w = 1:10, x = mod(16807 * x, 2147483647); end
    P = zeros(n, 680);
    for i = 1:n
        x = mod(16807 * x, 2147483647);
        P(i, :) = 2 * (x >= 1073741824) - 1;
    end
    KR_PROBES = P; KR_PROBE_POS = 0;
    addpath(shim);
    tr = trace_exp(A);
    rmpath(shim);
    out.(['C1_trace_exp_' names{g}]) = [tr KR_PROBE_POS];
end
if ~exist('golden_path', 'var'), golden_path = fullfile(root, 'tests', 'golden', 'reference_golden_c1_trace.json'); end
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
