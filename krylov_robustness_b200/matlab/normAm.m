function [c, mv] = normAm(A, m)
% Drop-in for functions/normAm.m.
[c, mv] = kr_mex('normAm', A, m);
end
