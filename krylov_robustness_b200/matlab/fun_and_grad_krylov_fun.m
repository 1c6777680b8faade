function [f, gr] = fun_and_grad_krylov_fun(X, A, Omega, fun, dfun, dfA, tol, it, debug, fun_M)
% Drop-in for functions/fun_and_grad_krylov_fun.m.
if ~ishermitian(A), error('FUN_AND_GRAD_KRYLOV_FCONNECTIVITY:: matrix A is not Hermitian'); end
[f, gr] = kr_mex('fun_and_grad', X(:), A, double(Omega), func2str(fun), func2str(dfun), dfA(:), tol, it);
end
