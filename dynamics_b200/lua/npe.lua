-- Drop-in replacement for the reference's src/lua/npe.lua: `M = require 'npe'` returns the same `model` table
-- (create, fp, fp_batch, bp, update_position, update_angle, cuda/float/clearState, fields network / criterion /
-- identitycriterion / theta.params / theta.grad_params / neighbor_masks / mp) with every tensor operation executed by
-- libnpe_cuda.so.  Reads the same globals (mp, config_args) the reference reads.
require 'nn'
require 'NPECuda'
require 'IdentityCriterion'
local npeffi = require 'npe_ffi'
local C, ffi = npeffi.C, npeffi.ffi

local model = {}
model.__index = model

function model.create(mp_, preload, model_path)               -- npe.lua:72-102
    local self = setmetatable({}, model)
    self.mp = mp_
    assert(self.mp.input_dim == self.mp.object_dim * self.mp.num_past)
    assert(self.mp.out_dim == self.mp.object_dim * self.mp.num_future)
    self.criterion = nn.MSECriterion(false)                    -- kept for checkpoint compatibility (data_utils.lua:210-230)
    self.identitycriterion = nn.IdentityCriterion()
    self.network = nn.NPECuda(self.mp)
    if preload then
        local checkpoint = torch.load(model_path)
        local src = checkpoint.model.theta and checkpoint.model.theta.params or checkpoint.model.network.weight
        self.network.weight:copy(src)                           -- same flat layout as the reference's getParameters()
        self.network:push()
    end
    self.theta = {}
    self.theta.params, self.theta.grad_params = self.network:getParameters()
    collectgarbage()
    return self
end

function model:cuda() end
function model:float() end
function model:clearState() self.network:clearState() end

local function flat_target(y) return y:contiguous():view(y:size(1), -1) end

function model:fp(params_, batch, sim)                         -- npe.lua:225-247
    if params_ ~= self.theta.params then self.theta.params:copy(params_) end
    self.network:push()
    self.theta.grad_params:zero()
    local this, context, y = unpack(batch)
    self.network.target = flat_target(y)
    local prediction = self.network:forward({this, context})
    local l = self.network.loss3
    return l[1], prediction, l[2], l[3]
end

function model:fp_batch(params_, batch, sim)                   -- npe.lua:250-284
    if params_ ~= self.theta.params then self.theta.params:copy(params_) end
    self.network:push()
    local this, context, y = unpack(batch)
    local B = this:size(1)
    local yy = flat_target(y)
    npeffi.check(self.network.handle, C.npe_batch_upload(self.network.handle, npeffi.fptr(this:contiguous()),
                 npeffi.fptr(context:contiguous()), npeffi.fptr(yy), B, context:size(2)))
    local pred, le = torch.FloatTensor(B, 22), torch.FloatTensor(B, 3)
    npeffi.check(self.network.handle, C.npe_forward_per_example(self.network.handle, npeffi.fptr(pred), npeffi.fptr(le)))
    return le:select(2, 1):clone(), pred, le:select(2, 2):clone(), le:select(2, 3):clone()
end

function model:bp(batch, prediction, sim)                      -- npe.lua:290-336
    self.theta.grad_params:zero()
    self.network:accGradParameters(nil, nil, 1)
    return self.theta.grad_params
end

-- the mask of the last batch, as a table of (B) tensors like select_neighbors returns (npe.lua:170-206)
function model:pull_neighbor_masks()
    local h = self.network.handle
    local F, K = C.npe_num_foci(h), C.npe_ctx_slots(h)
    local mask = torch.ByteTensor(F, K)
    npeffi.check(h, C.npe_build_pairs(h, nil, ffi.cast('uint8_t*', torch.data(mask)), nil))
    self.neighbor_masks = {}
    for c = 1, K do table.insert(self.neighbor_masks, mask:select(2, c):float()) end
    return self.neighbor_masks
end

-- Gradient with respect to the INPUTS (npe.lua:341-384).  Nothing on the north-star path calls it (main.lua / eval.lua -mode
-- sim never do; the reference version itself returns an undefined `d_input`): explicit error instead of a silent nil.
function model:bp_input(batch, prediction, sim)
    error('npe (libnpe_cuda): bp_input is not supported -- input gradients are not part of the training / rollout path')
end

-- simulate_all_postprocess helpers (eval.lua:170-173).  Host tensors of a few hundred floats: plain torch arithmetic with the
-- reference's operation order, (p * pnc + v * vnc) / pnc, each op rounded once, so positions stay bit-identical.
-- this (B, num_past, 22), pred (B, num_future, 22); returns a modified copy of pred.

-- Position of step k+1 = position of step k + velocity of step k, starting from the LAST PAST frame: the predicted
-- velocity of a step only moves the object one step later (npe.lua:386-419).
function model:update_position(this, pred)
    local si = config_args.si
    local pnc, vnc = config_args.position_normalize_constant, config_args.velocity_normalize_constant
    local out = pred:clone()
    local np_, nf = this:size(2), pred:size(2)
    local pos = this[{{}, {np_}, {si.px, si.py}}]:clone():mul(pnc)           -- (B,1,2) pixels
    local vel = this[{{}, {np_}, {si.vx, si.vy}}]:clone():mul(vnc)
    for k = 1, nf do
        pos = pos + vel                                                       -- integrate the PREVIOUS step's velocity
        out[{{}, {k}, {si.px, si.py}}] = pos / pnc
        vel = pred[{{}, {k}, {si.vx, si.vy}}]:clone():mul(vnc)
    end
    return out
end

-- Same recurrence for the angle, wrapped back into (-pi, pi] (npe.lua:422-458): a value above pi loses 2 pi, a value at or
-- below -pi gains 2 pi, applied to every step of the running sum at once like the reference does.
function model:update_angle(this, pred)
    local si = config_args.si
    local anc = config_args.angle_normalize_constant
    local out = pred:clone()
    local np_, nf = this:size(2), pred:size(2)
    local B = pred:size(1)
    local ang = torch.Tensor(B, nf + 1, 1):typeAs(pred)
    ang[{{}, {1}, {}}] = this[{{}, {np_}, {si.a}}]:clone():mul(anc)
    local av = this[{{}, {np_}, {si.av}}]:clone():mul(anc)
    for k = 1, nf do
        ang[{{}, {k + 1}, {}}] = ang[{{}, {k}, {}}] + av
        av = pred[{{}, {k}, {si.av}}]:clone():mul(anc)
    end
    local too_big = ang:gt(math.pi):typeAs(ang)
    local too_small = ang:le(-math.pi):typeAs(ang)
    ang = torch.add(ang, -2 * math.pi, too_big)
    ang = torch.add(ang, 2 * math.pi, too_small)
    out[{{}, {}, {si.a}}] = ang[{{}, {2, nf + 1}, {}}] / anc
    return out
end

return model
