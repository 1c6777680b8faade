function [M, mv, alpha, unA] = select_taylor_degree(A, b, m_max, p_max, prec, shift, bal, force_estm)
% Drop-in for functions/select_taylor_degree.m.
if nargin < 8, force_estm = false; end
if nargin < 6 || isempty(shift), shift = false; end
if nargin < 4 || isempty(p_max), p_max = 8; end
if nargin < 3 || isempty(m_max), m_max = 55; end
if p_max < 2 || m_max > 60 || m_max + 1 < p_max*(p_max - 1)
    error('>>> Invalid p_max or m_max.')
end
[M, mv, alpha, unA] = kr_mex('select_taylor_degree', A, size(b, 2), m_max, p_max, double(shift), double(force_estm));
end
