% dispmap_super -- B200 build of the base model class (surface of dispmap_super.m:3-329).
%
% Public properties and methods are those of the reference so example_global.m /
% example_simultaneous.m / example_ncc.m run unchanged; every array the fusion moves need
% (pairwise tables, q / qprim, energies, plane -> disparity) is produced on the GPU through
% sb_builders_mex, and the moves themselves go through rd.m / trws.m of this folder.
classdef dispmap_super < handle
	properties
		smoothness_kernel
		assignment;
		maxiter = 1000;      % TRW-S iterations / exhaustive binary fusion rounds
		max_relgap = 1e-4;   % TRW-S relative duality gap
		improve = false;     % run QPBO-I on unlabelled nodes
		device_loop = true;  % binary_fuse_until_convergence as ONE library call (fields stay on the GPU between moves)
		grid_native = false; % true: simultaneous_fusion through sb_grid_mex -- no L x E arrays (q, qprim) are ever built;
		                     % needed once they stop fitting (BASELINE configs 4-5)
	end
	properties (SetAccess = protected)
		sz;
		images;
		neighborhood;
		stored_energy;
		smooth_weights;
	end
	properties (Access = protected)
		dnorm = [0 1];       % [d_min d_step]: disparity normalisation (identity unless a subclass sets it)
	end
	methods
		function self = dispmap_super(images, kernel)
			self.images = images;
			dims = size(images{1});
			self.sz = dims(1:2);
			self.smoothness_kernel = kernel;
			construct_neighborhood(self);
			self.smooth_weights = ones(1, numel(self.neighborhood.ind1));
		end

		function set.max_relgap(self, v)
			if (v < 0), error('Maximum relative gap must be non-negative'); end
			self.max_relgap = v;
		end
		function set.improve(self, v)
			self.improve = logical(v);
		end
		function set.assignment(self, a)
			self.assignment = a;
			update_energy(self);
		end
		function set.smoothness_kernel(self, k)
			self.smoothness_kernel = k;
			update_energy(self);
		end
		function E = energy(self)
			E = self.stored_energy;
		end

		function [e, lb, num_unlabelled] = binary_fusion(self, proposal)
			% One QPBO fusion move between the current assignment and a proposal.
			if (~isequal(size(proposal), size(self.assignment)))
				error('Binary fusion: Proposals is of wrong size');
			end
			[E00, E01, E10, E11] = all_pairwise_costs(self, self.assignment, proposal);
			U0 = unary_cost(self, self.assignment);
			U1 = unary_cost(self, proposal);
			rd_options.improve = self.improve;
			[labelling, e, lb, num_unlabelled] = rd(U0, U1, E00, E01, E10, E11, connectivity(self), rd_options);
			take = (labelling == 1);
			fused = self.assignment;
			fused(:, take) = proposal(:, take);
			self.assignment = fused;
		end

		function number_of_iterations = binary_fuse_until_convergence(self, proposal_cell, show_steps)
			% Fuse proposals in a fixed-then-random order until a full round changes nothing.
			if nargin < 3, show_steps = false; end
			if ~iscell(proposal_cell), error('Input proposals should be given in cell array.'); end
			np = numel(proposal_cell);
			nrand = self.maxiter * 5;
			ids = [1:np, randi([1 np], nrand, 1)'];
			ids(diff(ids) == 0) = [];           % drop immediate repeats
			if self.device_loop && ~show_steps
				% the whole loop on device-resident fields; the visiting order is the one drawn above
				N = prod(self.sz);
				U = zeros(N, np);
				for l = 1:np
					if (~isequal(size(proposal_cell{l}), size(self.assignment)))
						error('Binary fusion: Proposals is of wrong size');
					end
					U(:, l) = unary_cost(self, proposal_cell{l});
				end
				[fused, E] = sb_builders_mex('fuse_until_convergence', self.sz, self.smoothness_kernel, cat(3, proposal_cell{:}), U, ...
					self.assignment, unary_cost(self, self.assignment), self.smooth_weights, self.tol, self.dnorm(1), self.dnorm(2), ...
					double(self.improve), self.maxiter, int32(ids));
				self.assignment = fused;
				number_of_iterations = numel(E);
				return;
			end
			E = energy(self);
			stale = false(np, 1);               % proposals tried since the last improvement
			for iter = 1:self.maxiter
				if iter > nrand, ids = [ids ids]; end %#ok<AGROW>
				k = iter + 1;                   % the reference starts at the second id (dispmap_super.m:116)
				if k > numel(ids), break; end
				if stale(ids(k)), continue; end
				binary_fusion(self, proposal_cell{ids(k)});
				E(end+1) = energy(self); %#ok<AGROW>
				if show_steps, display_current_dispmap(self); drawnow(); end
				if E(end-1) ~= E(end)
					stale(:) = false;
				else
					stale(ids(k)) = true;
				end
				if all(stale), break; end
			end
			number_of_iterations = numel(E);
		end

		function [e, lb, iterations] = simultaneous_fusion(self, proposal_cell)
			% TRW-S over all proposals plus the current assignment.
			if ~iscell(proposal_cell), error('Input proposals should be given in cell array.'); end
			proposal_cell{end+1} = self.assignment;
			L = numel(proposal_cell);
			N = prod(self.sz);
			unary = zeros(L, N);
			for l = 1:L
				unary(l, :) = unary_cost(self, proposal_cell{l});
			end
			stack = cat(3, proposal_cell{:});   % 4 x N x L
			options_struct.maxiter = self.maxiter;
			options_struct.max_relgap = self.max_relgap;
			if self.grid_native
				% the solver takes what this method holds (plane fields, unary slabs, weights) and recomputes the label
				% positions on the GPU: 45 bytes per label and pixel there, nothing of size L x E anywhere
				compile('sb_grid_mex.cpp', 'sb_grid_mex');
				[labels, e, lb, iterations] = sb_grid_mex(int32(self.smoothness_kernel), self.sz, stack, unary', ...
					self.smooth_weights(:), self.tol, self.dnorm, options_struct);
			else
				[q, qprim] = sb_builders_mex('fusion_positions', self.sz, stack, self.dnorm(1), self.dnorm(2));
				[labels, e, lb, iterations] = trws(int32(self.smoothness_kernel), unary, connectivity(self), q, qprim, ...
					self.smooth_weights(:), self.tol, options_struct);
			end
			fused = zeros(size(proposal_cell{1}));
			for l = 1:L
				pick = (labels == l);
				fused(:, pick) = proposal_cell{l}(:, pick);
			end
			self.assignment = fused;
		end

		function im = current_dispmap(self)
			im = reshape(disparitymap_from_assignment(self, self.assignment), self.sz);
		end
		function display_current_dispmap(self)
			imagesc(self.current_dispmap());
			colormap gray(256);
			title(sprintf('Solution energy: %g \n', self.energy()));
			axis equal;
		end
		function display(self)
			fprintf('Energy of current solution: %g. \n', self.energy());
			fprintf('Image pair size: (%d,%d) \n', self.sz(1), self.sz(2));
			fprintf('Settings: \n');
			fprintf('Smoothness kernel    : %d \n', self.smoothness_kernel)
			fprintf('Maximum iterations   : %d \n', self.maxiter)
			fprintf('Max relative gap     : %g \n', self.max_relgap);
			fprintf('RD-Improve			  : %d \n', int32(self.improve));
			display_current_dispmap(self);
		end
	end

	methods (Access = protected)
		function U = unary_cost(self, assignment) %#ok<INUSD,STOUT>
			error('Overload unary_cost \n');
		end
		function c = connectivity(self)
			c = uint32([self.neighborhood.ind1'; self.neighborhood.ind2']);
		end
		function [E00, E01, E10, E11] = all_pairwise_costs(self, assignment, proposals)
			% Pairwise tables of a fusion move, evaluated at the point of ind2 (GPU).
			if nargin < 3 || nargout < 2
				E00 = sb_builders_mex('pairwise_tables', self.sz, self.smoothness_kernel, assignment, [], ...
					self.smooth_weights, self.tol, self.dnorm(1), self.dnorm(2));
			else
				[E00, E01, E10, E11] = sb_builders_mex('pairwise_tables', self.sz, self.smoothness_kernel, assignment, ...
					proposals, self.smooth_weights, self.tol, self.dnorm(1), self.dnorm(2));
			end
		end
		function update_energy(self)
			if isempty(self.assignment) || isempty(self.smooth_weights)
				self.stored_energy = inf;
				return;
			end
			U = unary_cost(self, self.assignment);
			self.stored_energy = sb_builders_mex('energy', self.sz, self.smoothness_kernel, U, self.assignment, ...
				self.smooth_weights, self.tol, self.dnorm(1), self.dnorm(2));
		end
		function points = get_points(self)
			[xx, yy] = meshgrid(1:self.sz(2), 1:self.sz(1));
			points = [xx(:)'; yy(:)'];
		end
		function construct_neighborhood(self)
			% 4-connected grid, both directions of every neighbour pair: vertical pairs
			% (down, then up), then horizontal pairs (right, then left), column-major.
			nodenr = reshape(1:prod(self.sz), self.sz);
			vs = nodenr(1:end-1, :); vf = nodenr(2:end, :);
			hs = nodenr(:, 1:end-1); hf = nodenr(:, 2:end);
			self.neighborhood.ind1 = [vs(:); vf(:); hs(:); hf(:)];
			self.neighborhood.ind2 = [vf(:); vs(:); hf(:); hs(:)];
			self.neighborhood.nodenr = nodenr;
		end
		function set_disparity(self, disparity)
			a = zeros(4, prod(self.sz));
			a(3, :) = 1;
			a(4, :) = -disparity(:);
			self.assignment = a;
		end
		function init_assigments(self)
			a = zeros(4, prod(self.sz));
			a(3, :) = 1;
			self.assignment = a;
		end
		function disps = disparitymap_from_assignment(self, assignment, points)
			% plane [a b c d0] at point [x y] -> -(a x + b y + d0) / c, then (d - d_min) / d_step
			if nargin < 3, points = self.get_points(); end
			disps = sb_builders_mex('plane_disparity', assignment, points, self.dnorm(1), self.dnorm(2));
		end
	end
end
