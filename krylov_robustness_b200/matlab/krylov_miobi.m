function [edges, rob, A_new] = krylov_miobi(A, k, E, tol, it, poles, debug, miobi, rescale)
% Drop-in for functions/krylov_miobi.m: the candidate loop (:76-99) and the selection (:112-124) are ONE device
% call per round (kr_greedy_round); the edge update follows the reference (:127-137).
if ~issymmetric(A), error('KRYLOV_MIOBI:: Adjacency matrix should be symmetric'); end
if ~exist('tol', 'var'), tol = 1e-12; end
if ~exist('it', 'var'), it = min(100, size(A, 1)); end
if ~exist('miobi', 'var'), miobi = 'break'; end
if ~exist('rescale', 'var'), rescale = 1; end
if ~exist('E', 'var') || isempty(E)
    [t1, t2] = find(A); ind = find(t1 >= t2); E = [t1(ind), t2(ind)];
end
if ~strcmp(miobi, 'break') && ~strcmp(miobi, 'make'), error('KRYLOV_MIOBI:: not supported option for miobi'); end
if strcmp(miobi, 'break') && nnz(A) < 2*k
    error('KRYLOV_MIOBI:: edges to be removed are more than edges in the network')
end
sgn = 1; if strcmp(miobi, 'break'), sgn = -1; end
rob = 0; edges = zeros(0, 2);
hA = kr_mex('matrix_create', A);                    % A goes to the device ONCE; the rounds edit it in place
freeA = onCleanup(@() kr_mex('matrix_free', hA));
for j = 1:min(k, size(E, 1))
    % scores of all candidates + the first-wins selection of :112-124 in ONE device call (kr_greedy_round); with
    % KR_GREEDY_SCREEN=1 in the environment candidates that cannot win are ruled out by shared node bases and only the
    % contenders take the block Lanczos path - the selected edge and its value are the same either way.
    % self loops are not rescaled (:88-94)
    [best, bestval] = kr_mex('greedy_round', hA, double(E), sgn/rescale, tol, it, 'exp', sgn, double(sgn > 0), -1);
    mx = [best bestval];
    chosen = E(mx(1), :);
    E = E([1:mx(1)-1, mx(1)+1:end], :);
    A(chosen(1), chosen(2)) = (sgn > 0); A(chosen(2), chosen(1)) = (sgn > 0);
    kr_mex('matrix_set_edges', hA, chosen(1), chosen(2), double(sgn > 0));      % :127-135 on the device copy
    edges = [edges; chosen]; rob = rob + mx(2);
end
A_new = A;
end
