function [V, K, H, params, lucky] = arnoldi_krylov(varargin)
% Drop-in for functions/arnoldi_krylov.m: (A,b) starts, (V,K,H,params) extends.  A may be an operator
% struct from kr_operator(A) (functions/arnoldi_krylov.m:34,83-84); see lanczos_krylov.m.
if nargin ~= 2 && nargin ~= 4, error('Called with the wrong number of arguments'); end
if nargin == 2
    A = varargin{1}; b = varargin{2};
    if isstruct(A)
        if ~isfield(A, 'matrix')
            error('krylov_b200:operator', 'operator structs must come from kr_operator(A): a MATLAB function handle cannot run on the device');
        end
        S = A.matrix;
    else
        if size(A, 1) ~= size(A, 2), error('The matrix A should be square'); end
        if size(A, 2) ~= size(b, 1), error('The block vector b has wrong number of rows'); end
        S = A;
    end
    [V, H, K, last, lucky, h] = kr_mex('krylov_start', S, full(b), 1);
    params = struct('last', last, 'A', A, 'handle', h, 'n', size(S, 1), ...
                    'cleanup', onCleanup(@() kr_mex('krylov_free', h)));
else
    params = varargin{4};
    [V, H, K, last, lucky, ~] = kr_mex('krylov_extend', params.handle, params.n);
    params.last = last;
end
end
