function [Xm, iter, lucky, Um] = fun_update(A, U, B, fun, tol, it, debug)
% Drop-in for functions/fun_update.m; nargout == 4 selects Arnoldi + basis (:69,:77).
if ~exist('tol', 'var') || isempty(tol), tol = 1e-12; end
if ~exist('it', 'var') || isempty(it), it = min(100, size(A, 1)); end
if nargout >= 4
    [Xm, iter, lucky, Um] = kr_mex('fun_update', A, full(U), full(B), func2str(fun), tol, it, 1);
else
    [Xm, iter, lucky] = kr_mex('fun_update', A, full(U), full(B), func2str(fun), tol, it, 0);
end
if lucky, warning('FUN_UPDATE:: Detected lucky breakdown'); end
if iter == it, warning('FUN_UPDATE:: Reached maximum number of iterations'); end
end
