function Hes = hessianfcn_fun(X, A, Omega, f, tol, it)
% Drop-in for functions/hessianfcn_fun.m.
n = size(A, 1);
XX = sparse(Omega(:, 1), Omega(:, 2), X(:), n, n);
Atilde = A + XX + XX';
Hes = kr_mex('hessian', Atilde, double(Omega), func2str(f), tol, it);
end
