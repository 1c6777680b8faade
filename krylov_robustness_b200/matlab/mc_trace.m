function [tr_new, res, it] = mc_trace(Afun, n, tol, maxit, isAreal, debug)
% Drop-in for functions/mc_trace.m in parity mode.  Afun is a matrix, or the struct
% struct('expmv_of', A) that trace_exp.m builds in place of @(x) expmv(1,A,x,[],'double').
% Probes are drawn here exactly as the reference does (:43-44), K = ceil(maxit/30) pairs up front.
if ~exist('tol', 'var'), tol = 1e-3; end
if ~exist('maxit', 'var'), maxit = 10; end
K = ceil(maxit/30);
probes = zeros(n, 20*K);
for i = 1:K
    probes(:, 20*(i-1)+(1:10)) = sign(randn(n, 10));
    probes(:, 20*(i-1)+(11:20)) = sign(randn(n, 10));
end
if isstruct(Afun)
    [tr_new, res, it] = kr_mex('mc_trace', Afun.expmv_of, 1, tol, maxit, probes);
else
    [tr_new, res, it] = kr_mex('mc_trace', Afun, 0, tol, maxit, probes);
end
end
