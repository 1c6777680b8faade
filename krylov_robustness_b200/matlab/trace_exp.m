function tr = trace_exp(A)
% Drop-in for functions/trace_exp.m.
tr = mc_trace(struct('expmv_of', A), size(A, 1), 1e-4, 1000, 1);
end
