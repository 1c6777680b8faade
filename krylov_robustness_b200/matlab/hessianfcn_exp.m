function Hes = hessianfcn_exp(X, A, Omega, tol, it)
% Drop-in for functions/hessianfcn_exp.m: Atilde is assembled here (:4-7), the Frechet factors and the
% bilinear forms (functions/multiple_frechet_eval.m) run on the device.
n = size(A, 1);
XX = sparse(Omega(:, 1), Omega(:, 2), X(:), n, n);
Atilde = A + XX + XX';
Hes = kr_mex('hessian', Atilde, double(Omega), 'exp', tol, it);
end
