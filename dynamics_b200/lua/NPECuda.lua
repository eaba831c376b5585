-- nn.NPECuda: the whole NPE network (encoder -> 5 pairwise layers -> sum -> decoder) as ONE nn.Module backed by
-- libnpe_cuda.so.  It replaces the nn.Sequential built by init_network (reference src/lua/npe.lua:14-61) and keeps the
-- surface main.lua / eval.lua / checkpointing rely on:
--   updateOutput / updateGradInput / accGradParameters, parameters() -> {flat}, {gradflat}  so that getParameters()
--   returns the same 28 222-float layout, clearState, type/float/cuda (no-ops: the weights live on the GPU behind the
--   handle and are mirrored into `self.weight` for torch.save), read/write for .t7 checkpoints.
require 'nn'
local npeffi = require 'npe_ffi'
local C, ffi = npeffi.C, npeffi.ffi

local NPECuda, parent = torch.class('nn.NPECuda', 'nn.Module')

function NPECuda:__init(mp)
    parent.__init(self)
    self.mp = mp
    local cfg = ffi.new('npe_config[1]')
    C.npe_default_config(cfg)
    cfg[0].device = mp.gpu_device or 0
    cfg[0].rnn_dim = mp.rnn_dim; cfg[0].layers = mp.layers
    cfg[0].num_past = mp.num_past; cfg[0].num_future = mp.num_future; cfg[0].object_dim = mp.object_dim
    cfg[0].nbrhd = mp.nbrhd and 1 or 0; cfg[0].nbrhdsize = mp.nbrhdsize
    cfg[0].vlambda = mp.vlambda; cfg[0].lambda = mp.lambda
    local hp = ffi.new('npe_handle*[1]')
    local rc = C.npe_create(cfg, hp)
    if rc ~= 0 then error('npe_create: ' .. ffi.string(C.npe_last_error(nil))) end
    self.handle = ffi.gc(hp[0], C.npe_destroy)
    local n = C.npe_param_count(self.handle)
    self.weight = torch.FloatTensor(n)        -- flat parameters, reference getParameters() layout
    self.gradWeight = torch.FloatTensor(n):zero()
    self:reset()
    self.output = torch.FloatTensor()
end

-- nn.Linear:reset() per layer: U(-1/sqrt(in), 1/sqrt(in)); the five pairwise layers start identical (npe.lua:39)
function NPECuda:reset()
    local w, off = self.weight, 0
    local function fill(n_out, n_in, bias, src)
        local stdv = 1 / math.sqrt(n_in)
        local blk = w:narrow(1, off + 1, n_out * n_in)
        if src then blk:copy(src) else blk:uniform(-stdv, stdv) end
        off = off + n_out * n_in
        if bias then w:narrow(1, off + 1, n_out):uniform(-stdv, stdv); off = off + n_out end
        return blk
    end
    fill(25, 44, false); fill(25, 44, false)
    local first = fill(50, 50, false)
    for _ = 2, 5 do fill(50, 50, false, first) end
    fill(50, 94, true); fill(50, 50, true); fill(50, 50, true); fill(50, 50, true); fill(22, 50, true)
    assert(off == w:nElement())
    self:push()
end

function NPECuda:push()   -- host mirror -> device
    npeffi.check(self.handle, C.npe_params_set(self.handle, npeffi.fptr(self.weight), self.weight:nElement()))
end
function NPECuda:pull()   -- device -> host mirror (before torch.save / after optimiser steps)
    npeffi.check(self.handle, C.npe_params_get(self.handle, npeffi.fptr(self.weight), self.weight:nElement()))
end

function NPECuda:parameters() return {self.weight}, {self.gradWeight} end

-- input = {this (B,2,22), context (B,C,2,22)} (the untouched batch tensors; pairing, masking and trimming happen on
-- the GPU), optional self.target (B,22) for the fused loss.
function NPECuda:updateOutput(input)
    local this, context = input[1]:contiguous(), input[2]:contiguous()
    local B, Cn = this:size(1), context:size(2)
    local y = self.target and self.target:contiguous() or nil
    npeffi.check(self.handle, C.npe_batch_upload(self.handle, npeffi.fptr(this), npeffi.fptr(context),
                                                 y and npeffi.fptr(y) or nil, B, Cn))
    self.output:resize(B, 22)
    self.loss3 = self.loss3 or torch.FloatTensor(3)
    npeffi.check(self.handle, C.npe_forward(self.handle, npeffi.fptr(self.output), y and npeffi.fptr(self.loss3) or nil))
    return self.output
end

-- the reference never needs gradients w.r.t. the raw states on the training path (npe.lua:328-330 discards them)
function NPECuda:updateGradInput(input, gradOutput)
    self.gradInput = self.gradInput or {}
    return self.gradInput
end

-- gradOutput is implied by the fused loss (d_vel, d_ang_vel of npe.lua:301-320); scale is folded into vlambda/lambda
function NPECuda:accGradParameters(input, gradOutput, scale)
    npeffi.check(self.handle, C.npe_backward(self.handle, self.divisor_batch or 0))
    npeffi.check(self.handle, C.npe_grads_get(self.handle, npeffi.fptr(self.gradWeight), self.gradWeight:nElement()))
end

function NPECuda:clearState() self.output:set(); return self end
function NPECuda:type(t) return self end        -- weights live on the GPU; host mirrors stay FloatTensor
function NPECuda:float() return self end
function NPECuda:cuda() return self end

-- .t7: the checkpoint stores the flat vector; a reference checkpoint's model.theta.params loads into self.weight
function NPECuda:write(file)
    self:pull()
    file:writeObject({weight = self.weight, mp = self.mp})
end
function NPECuda:read(file)
    local t = file:readObject()
    self.__index = nil
    NPECuda.__init(self, t.mp)
    self.weight:copy(t.weight)
    self:push()
end
