function [f, s, m, mv, mvd, unA] = expmv(t, A, b, M, prec, shift, bal, full_term, prnt)
% Drop-in for functions/expmv.m (double precision, bal = false).
if nargin < 8 || isempty(full_term), full_term = false; end
if nargin < 7 || isempty(bal), bal = false; end
if nargin < 6 || isempty(shift), shift = true; end
if nargin < 4, M = []; end
if bal, error('expmv (B200): balancing is not provided'); end
[f, s, m, mv, mvd, unA] = kr_mex('expmv', t, A, full(b), M, double(shift), double(full_term));
end
