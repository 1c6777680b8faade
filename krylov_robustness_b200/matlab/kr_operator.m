function op = kr_operator(A)
% op = kr_operator(A): the operator struct of functions/lanczos_krylov.m:32,78-79 / arnoldi_krylov.m:34,83-84
% for a device-resident matrix.  op.multiply(alpha, beta, w) follows the reference's call
% A.multiply(1.0, 0.0, w) (y = alpha*A*w; the beta term multiplies the zero matrix) and runs the SpMM on the
% device, so reference code that only knows the struct protocol keeps working; lanczos_krylov / arnoldi_krylov
% of this package recognise the field `matrix` and keep the whole Krylov state on the device.
if size(A, 1) ~= size(A, 2), error('The matrix A should be square'); end
op = struct('matrix', A, 'multiply', @(alpha, beta, w) alpha * kr_mex('spmm', A, full(w)));
end
