function [V, H, params, lucky] = lanczos_krylov(varargin)
% Drop-in for functions/lanczos_krylov.m: (A,b) starts, (V,H,params) extends.  The Krylov state lives
% on the device behind params.handle; V (two-block window) and H are copied out on every call.
if nargin ~= 2 && nargin ~= 3, error('Called with the wrong number of arguments'); end
if nargin == 2
    A = varargin{1}; b = varargin{2};
    if size(A, 1) ~= size(A, 2), error('The matrix A should be square'); end
    if size(A, 2) ~= size(b, 1), error('The block vector b has wrong number of rows'); end
    [V, H, ~, last, lucky, h] = kr_mex('krylov_start', A, full(b), 0);
    params = struct('last', last, 'A', A, 'handle', h, 'n', size(A, 1));
else
    params = varargin{3};
    [V, H, ~, last, lucky, ~] = kr_mex('krylov_extend', params.handle, params.n);
    params.last = last;
end
end
