function [f, gr] = fun_and_grad_krylov_exp(X, A, Omega, eA, tol, it, debug)
% Drop-in for functions/fun_and_grad_krylov_exp.m.
if ~ishermitian(A), error('FUN_AND_GRAD_KRYLOV:: matrix A is not Hermitian'); end
[f, gr] = kr_mex('fun_and_grad', X(:), A, double(Omega), 'exp', 'exp', eA(:), tol, it);
end
