function [V, H, params, lucky] = lanczos_krylov(varargin)
% Drop-in for functions/lanczos_krylov.m: (A,b) starts, (V,H,params) extends.  The Krylov state lives
% on the device behind params.handle; V (two-block window) and H are copied out on every call.
% A may be the operator struct of the reference (functions/lanczos_krylov.m:32,78-79): a struct made by
% kr_operator(A) carries the sparse matrix in its field `matrix` and runs on the device with the
% reference's validation skipped, exactly as the reference skips it for structs.  The handle is released
% when the last copy of params goes out of scope (params.cleanup) or at mexAtExit.
if nargin ~= 2 && nargin ~= 3, error('Called with the wrong number of arguments'); end
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
    [V, H, ~, last, lucky, h] = kr_mex('krylov_start', S, full(b), 0);
    params = struct('last', last, 'A', A, 'handle', h, 'n', size(S, 1), ...
                    'cleanup', onCleanup(@() kr_mex('krylov_free', h)));
else
    params = varargin{3};
    [V, H, ~, last, lucky, ~] = kr_mex('krylov_extend', params.handle, params.n);
    params.last = last;
end
end
