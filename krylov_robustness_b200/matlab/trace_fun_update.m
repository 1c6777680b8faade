function [Xm, iter, lucky] = trace_fun_update(A, U, B, tol, it, debug, fun)
% Drop-in for functions/trace_fun_update.m (same signature and defaults, :21-35); the arithmetic runs
% on the B200 through kr_mex -> kr_trace_fun_update (include/krylov_b200.h).
if ~exist('tol', 'var') || isempty(tol), tol = 1e-12; end
if ~exist('it', 'var') || isempty(it), it = min(100, size(A, 1)); end
if ~exist('fun', 'var') || isempty(fun), fun = @exp; end
[Xm, iter, lucky] = kr_mex('trace_fun_update', A, full(U), full(B), tol, it, func2str(fun));
if iter == it
    warning('TRACE_FUN_UPDATE:: Reached maximum number of iterations')
end
end
