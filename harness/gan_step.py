"""End-to-end measurement harness: one full SBA-GAN G+D adversarial
training step around the fused hot path, on synthetic CUB-shaped data with random-init weights.

This is MEASUREMENT INFRASTRUCTURE, not part of the product: the networks below are a compact
from-scratch restatement of the reference's architecture at bird_style.yml dimensions
(AttnGAN2/code/model_bert.py:304-594 generator with AdaIN + mapping network,
AttnGAN2/code/model.py:540-674 discriminators, model.py:162-267 Inception-v3 image encoder,
model_bert.py:161-189 BERT caption encoder) so that the step body of
AttnGAN2/code/trainer_bert.py:251-304 can run on a box where /root/reference does not exist.
Everything except the attention module and words_loss is stock torch / cuDNN, exactly as in the
reference; `attention="fused"` plugs in sba_gan_b200.GlobalAttentionGeneral / words_loss,
`attention="eager"` runs the reference's eager op sequence for the same two ops on the GPU.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

NGF, NDF, NEF, NCF, ZDIM, WDIM, RNUM, WORDS = 32, 64, 256, 100, 100, 256, 2, 20   # cfg/bird_style.yml, config.py
GAMMA1, GAMMA2, GAMMA3, LAMBDA = 4.0, 5.0, 10.0, 5.0                                  # cfg/bird_style.yml:28-31


# ------------------------------------------------------------------ eager reference ops (GPU)
class EagerAttention(nn.Module):
    """The reference's op sequence (GlobalAttention.py:82-121) in stock torch ops."""

    def __init__(self, idf, cdf):
        super().__init__()
        self.conv_context = nn.Conv2d(cdf, idf, 1, bias=False)
        self.mask = None

    def applyMask(self, mask):
        self.mask = mask

    def forward(self, x, context):
        B, idf, ih, iw = x.shape
        Q, L = ih * iw, context.shape[2]
        tgt = x.view(B, idf, Q).transpose(1, 2).contiguous()
        src = self.conv_context(context.unsqueeze(3)).squeeze(3)
        s = torch.bmm(tgt, src).view(B * Q, L)
        if self.mask is not None:
            s.data.masked_fill_(self.mask.repeat(Q, 1), float("-inf"))
        p = torch.softmax(s, dim=1).view(B, Q, L).transpose(1, 2).contiguous()
        c = torch.bmm(src, p)
        return c.view(B, idf, ih, iw), p.view(B, L, ih, iw)


def eager_words_loss(img, words, labels, cap_lens, class_ids, B):
    """miscc/losses.py:62-132 restated with stock torch ops (per-caption loop, as the reference)."""
    lens = cap_lens.tolist()
    R = img.shape[2] * img.shape[3]
    ctx = img.view(B, NEF, R)
    ctxT = ctx.transpose(1, 2).contiguous()
    sims = []
    for i in range(B):
        T = lens[i]
        w = words[i, :, :T].unsqueeze(0).expand(B, NEF, T)
        a = torch.softmax(torch.bmm(ctxT, w), dim=2)                       # over words
        a = torch.softmax(a.transpose(1, 2) * GAMMA1, dim=2)               # over regions
        wc = torch.bmm(ctx, a.transpose(1, 2))                             # B x nef x T
        cos = F.cosine_similarity(w, wc, dim=1, eps=1e-8)                  # B x T
        sims.append(torch.log(torch.exp(cos * GAMMA2).sum(1, keepdim=True)))
    sim = torch.cat(sims, 1) * GAMMA3
    if class_ids is not None:
        same = class_ids[:, None] == class_ids[None, :]
        same.fill_diagonal_(False)
        sim = sim.masked_fill(same, float("-inf"))
    return F.cross_entropy(sim, labels), F.cross_entropy(sim.t(), labels), None


# ------------------------------------------------------------------ building blocks
def glu(x):
    a, b = x.chunk(2, 1)
    return a * torch.sigmoid(b)


class GLU(nn.Module):
    def forward(self, x):
        return glu(x)


def conv3(i, o):
    return nn.Conv2d(i, o, 3, 1, 1, bias=False)


def up_block(i, o):
    return nn.Sequential(nn.Upsample(scale_factor=2, mode="nearest"), conv3(i, 2 * o), nn.BatchNorm2d(2 * o), GLU())


class ResBlock(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.block = nn.Sequential(conv3(c, 2 * c), nn.BatchNorm2d(2 * c), GLU(), conv3(c, c), nn.BatchNorm2d(c))

    def forward(self, x):
        return self.block(x) + x


class AdaIN(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.norm = nn.InstanceNorm2d(c)
        self.style = nn.Linear(WDIM, 2 * c)

    def forward(self, h, w):
        g, b = self.style(w)[:, :, None, None].chunk(2, 1)
        return (g + 1.0) * self.norm(h) + b


class NextStage(nn.Module):
    def __init__(self, att_cls, fuse_stage=False):
        super().__init__()
        self.fuse_stage = fuse_stage            # attention + AdaIN + cat as one operator (sba_gan_b200.stage)
        self.att = att_cls(NGF, NEF)
        self.adain2 = AdaIN(NGF)
        self.residual = nn.Sequential(*[ResBlock(2 * NGF) for _ in range(RNUM)])
        self.upsample = up_block(2 * NGF, NGF)

    def forward(self, h, w_code, words, mask):
        self.att.applyMask(mask)
        if self.fuse_stage:
            from sba_gan_b200.stage import next_stage_attention
            h_c, att = next_stage_attention(self.att, self.adain2, h, w_code, words)
            return self.upsample(self.residual(h_c)), att
        c_code, att = self.att(h, words)
        h = self.adain2(h, w_code)
        return self.upsample(self.residual(torch.cat((h, c_code), 1))), att


class Generator(nn.Module):
    def __init__(self, att_cls, fuse_stage=False):
        super().__init__()
        self.ca_fc = nn.Linear(NEF, 4 * NCF)
        self.mapping = nn.Sequential(nn.Linear(ZDIM, WDIM, bias=False), *[nn.Linear(WDIM, WDIM, bias=False) for _ in range(7)])
        g = 16 * NGF
        self.fc = nn.Sequential(nn.Linear(NCF, g * 4 * 4 * 2, bias=False), nn.BatchNorm1d(g * 4 * 4 * 2), GLU())
        self.ups = nn.Sequential(up_block(g, g // 2), up_block(g // 2, g // 4), up_block(g // 4, g // 8), up_block(g // 8, g // 16))
        self.h_net2, self.h_net3 = NextStage(att_cls, fuse_stage), NextStage(att_cls, fuse_stage)
        self.img = nn.ModuleList([nn.Sequential(conv3(NGF, 3), nn.Tanh()) for _ in range(3)])

    def forward(self, z, sent, words, mask):
        x = glu(self.ca_fc(sent))
        mu, logvar = x[:, :NCF], x[:, NCF:]
        c = torch.randn_like(mu) * torch.exp(0.5 * logvar) + mu
        w = self.mapping(z)
        h1 = self.ups(self.fc(c).view(-1, 16 * NGF, 4, 4))
        h2, a1 = self.h_net2(h1, w, words, mask)
        h3, a2 = self.h_net3(h2, w, words, mask)
        return [self.img[0](h1), self.img[1](h2), self.img[2](h3)], [a1, a2], mu, logvar


def down_block(i, o):
    return nn.Sequential(nn.Conv2d(i, o, 4, 2, 1, bias=False), nn.BatchNorm2d(o), nn.LeakyReLU(0.2, inplace=True))


def leaky3(i, o):
    return nn.Sequential(conv3(i, o), nn.BatchNorm2d(o), nn.LeakyReLU(0.2, inplace=True))


class Logits(nn.Module):
    def __init__(self, cond):
        super().__init__()
        self.joint = leaky3(8 * NDF + NEF, 8 * NDF) if cond else None
        self.out = nn.Sequential(nn.Conv2d(8 * NDF, 1, 4, 4), nn.Sigmoid())

    def forward(self, h, c=None):
        if self.joint is not None and c is not None:
            h = self.joint(torch.cat((h, c.view(-1, NEF, 1, 1).repeat(1, 1, 4, 4)), 1))
        return self.out(h).view(-1)


class Discriminator(nn.Module):
    def __init__(self, size):
        super().__init__()
        n = NDF
        layers = [nn.Conv2d(3, n, 4, 2, 1, bias=False), nn.LeakyReLU(0.2, inplace=True), down_block(n, 2 * n),
                  down_block(2 * n, 4 * n), down_block(4 * n, 8 * n)]
        if size >= 128:
            layers.append(down_block(8 * n, 16 * n))
        if size >= 256:
            layers += [down_block(16 * n, 32 * n), leaky3(32 * n, 16 * n)]
        if size >= 128:
            layers.append(leaky3(16 * n, 8 * n))
        self.encode = nn.Sequential(*layers)
        self.UNCOND_DNET, self.COND_DNET = Logits(False), Logits(True)

    def forward(self, x):
        return self.encode(x)


class ImageEncoder(nn.Module):
    """Inception-v3 up to Mixed_6e -> 17x17x768 -> 1x1 conv to nef; pooled 2048 -> nef (model.py:162-267)."""

    def __init__(self):
        super().__init__()
        import torchvision
        self.net = torchvision.models.inception_v3(weights=None, aux_logits=True, init_weights=False)
        self.emb_features = nn.Conv2d(768, NEF, 1, bias=False)
        self.emb_cnn_code = nn.Linear(2048, NEF)
        for p in self.net.parameters():
            p.requires_grad_(False)

    def forward(self, x):
        n = self.net
        x = F.interpolate(x, size=(299, 299), mode="bilinear", align_corners=False)
        x = n.Conv2d_2b_3x3(n.Conv2d_2a_3x3(n.Conv2d_1a_3x3(x)))
        x = F.max_pool2d(x, 3, 2)
        x = n.Conv2d_4a_3x3(n.Conv2d_3b_1x1(x))
        x = F.max_pool2d(x, 3, 2)
        x = n.Mixed_5d(n.Mixed_5c(n.Mixed_5b(x)))
        x = n.Mixed_6e(n.Mixed_6d(n.Mixed_6c(n.Mixed_6b(n.Mixed_6a(x)))))
        feat = self.emb_features(x)                                 # B x nef x 17 x 17
        x = n.Mixed_7c(n.Mixed_7b(n.Mixed_7a(x)))
        code = self.emb_cnn_code(F.adaptive_avg_pool2d(x, 1).flatten(1))
        return feat, code


class TextEncoder(nn.Module):
    """Frozen bert-base (random init here) -> 1x1 conv 768 -> nef + tanh (model_bert.py:161-189)."""

    def __init__(self):
        super().__init__()
        from transformers import BertConfig, BertModel
        self.bert = BertModel(BertConfig())
        self.conv_text = nn.Conv1d(768, NEF, 1)
        self.fc_sent = nn.Linear(768, NEF)

    @torch.no_grad()
    def forward(self, captions):
        out = self.bert(input_ids=captions, attention_mask=(captions != 0).long())
        words = torch.tanh(self.conv_text(out.last_hidden_state.transpose(1, 2)))
        return words, torch.tanh(self.fc_sent(out.pooler_output))


# ------------------------------------------------------------------ losses (miscc/losses.py)
def bce(p, t):
    """nn.BCELoss; evaluated in fp32 outside autocast (BCE on probabilities is not autocast-safe)."""
    with torch.autocast("cuda", enabled=False):
        return F.binary_cross_entropy(p.float(), t)


def sent_loss(cnn_code, sent, labels, class_ids):
    a = cnn_code / cnn_code.norm(dim=1, keepdim=True).clamp_min(1e-8)
    b = sent / sent.norm(dim=1, keepdim=True).clamp_min(1e-8)
    s = a @ b.t() * GAMMA3
    if class_ids is not None:
        same = class_ids[:, None] == class_ids[None, :]
        same.fill_diagonal_(False)
        s = s.masked_fill(same, float("-inf"))
    return F.cross_entropy(s, labels), F.cross_entropy(s.t(), labels)


def d_loss(netD, real, fake, cond, ones, zeros):
    rf, ff = netD(real), netD(fake.detach())
    c_real, c_fake = bce(netD.COND_DNET(rf, cond), ones), bce(netD.COND_DNET(ff, cond), zeros)
    c_wrong = bce(netD.COND_DNET(rf[:-1], cond[1:]), zeros[1:])
    u_real, u_fake = bce(netD.UNCOND_DNET(rf), ones), bce(netD.UNCOND_DNET(ff), zeros)
    return (u_real + c_real) / 2.0 + (u_fake + c_fake + c_wrong) / 3.0


class Trainer:
    """Step body of trainer_bert.py:251-304 on synthetic data; one instance per rank."""

    def __init__(self, batch, device, attention="fused", world=1, seed=0):
        torch.manual_seed(seed)
        if attention in ("fused", "fused_stage"):
            from sba_gan_b200 import GlobalAttentionGeneral as att_cls, words_loss as wl, sent_loss as sl
            self.words_loss = lambda img, w, lab, lens, cls, B: wl(img, w, lab, lens, cls, B, GAMMA1, GAMMA2, GAMMA3,
                                                                   att_maps=False)
            self.sent_loss = lambda code, sent, lab, cls: sl(code, sent, lab, cls, self.B, gamma3=GAMMA3)
        else:
            att_cls, self.words_loss, self.sent_loss = EagerAttention, eager_words_loss, sent_loss
        self.B, self.dev, self.world = batch, device, world
        self.G = Generator(att_cls, fuse_stage=(attention == "fused_stage")).to(device)
        self.Ds = [Discriminator(s).to(device) for s in (64, 128, 256)]
        self.img_enc = ImageEncoder().to(device).eval()
        self.txt_enc = TextEncoder().to(device).eval()
        for p in self.img_enc.parameters():
            p.requires_grad_(False)
        self.optG = torch.optim.Adam(self.G.parameters(), lr=2e-4, betas=(0.5, 0.999))
        self.optD = [torch.optim.Adam(d.parameters(), lr=2e-4, betas=(0.5, 0.999)) for d in self.Ds]
        self.avg = [p.detach().clone() for p in self.G.parameters()]
        g = torch.Generator().manual_seed(seed + 1)
        self.caps = torch.zeros(batch, WORDS, dtype=torch.long)
        self.lens = torch.sort(torch.randint(5, WORDS + 1, (batch,), generator=g), descending=True).values
        for i, n in enumerate(self.lens.tolist()):
            self.caps[i, :n] = torch.randint(1000, 20000, (n,), generator=g)
        self.caps, self.lens = self.caps.to(device), self.lens.to(device)
        self.cls = torch.randint(1, 201, (batch,), generator=g).to(device)
        self.real = [torch.randn(batch, 3, s, s, generator=g).to(device) for s in (64, 128, 256)]
        self.ones, self.zeros = torch.ones(batch, device=device), torch.zeros(batch, device=device)
        self.labels = torch.arange(batch, device=device)

    def _reduce(self, params):
        if self.world > 1:
            from sba_gan_b200.parallel import allreduce_gradients
            allreduce_gradients(params)

    def step(self):
        B = self.B
        words, sent = self.txt_enc(self.caps)                       # frozen, detached (trainer_bert.py:256-257)
        mask = (self.caps == 0)[:, :words.shape[2]]
        noise = torch.randn(B, ZDIM, device=self.dev)
        fake, _, mu, logvar = self.G(noise, sent, words, mask)
        errD = 0.0
        for D, opt, real, f in zip(self.Ds, self.optD, self.real, fake):
            D.zero_grad(set_to_none=True)
            e = d_loss(D, real, f, sent, self.ones, self.zeros)
            e.backward()
            self._reduce(D.parameters())
            opt.step()
            errD = errD + e.detach()
        self.G.zero_grad(set_to_none=True)
        errG = 0.0
        for D, f in zip(self.Ds, fake):
            h = D(f)
            errG = errG + bce(D.UNCOND_DNET(h), self.ones) + bce(D.COND_DNET(h, sent), self.ones)
        feat, code = self.img_enc(fake[-1])
        w0, w1, _ = self.words_loss(feat, words, self.labels, self.lens, self.cls, B)
        s0, s1 = self.sent_loss(code, sent, self.labels, self.cls)
        kl = -0.5 * torch.mean(1 + logvar - mu.pow(2) - logvar.exp())
        errG = errG + (w0 + w1 + s0 + s1) * LAMBDA + kl
        errG.backward()
        self._reduce(self.G.parameters())
        self.optG.step()
        with torch.no_grad():
            torch._foreach_mul_(self.avg, 0.999)
            torch._foreach_add_(self.avg, [p.detach() for p in self.G.parameters()], alpha=0.001)
        return errD, errG.detach()
