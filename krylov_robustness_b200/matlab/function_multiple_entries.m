function [X, iter] = function_multiple_entries(A, omega, f, tol, it, poles, debug)
% Drop-in for functions/function_multiple_entries.m (poles = inf only, :92).
if ~exist('tol', 'var') || isempty(tol), tol = 1e-12; end
if ~exist('it', 'var') || isempty(it), it = min(100, size(A, 1)); end
if exist('poles', 'var') && ~(isscalar(poles) && poles == inf)
    error('FUNCTION_MULTIPLE_ENTRIES::Unsupported rational Krylov yet')
end
[X, iter] = kr_mex('function_multiple_entries', A, double(omega), func2str(f), tol, it);
if iter == it
    warning('FUNCTION_MULTIPLE_ENTRIES:: Reached maximum number of iterations')
end
end
