# Drop-in replacement of genomicMateSelectR::predCrossVars (R/predCrossVar.R:24-78 of the reference)
# backed by libgms_b200.so through .Call. Same arguments, same return tibble.
# NOT executable in this repository's build image (no R there): this R code has never been run. The .Call shim
# under it IS executed on the GPU box against a stand-in for R's C API (r-package/tests/fakeR), and the Python module
# genomicmateselectr_b200/predcrossvar.py is the executable twin of this file and is what the tests run.
#
# NAMESPACE additions:  useDynLib(genomicMateSelectRb200, .registration = TRUE)

# --- stage (1), lazily: keep the map instead of the m x m matrix -------------------------------
# genmap2recombfreq_b200(m, nChr) returns an object for which `1-(2*x)` (what users write,
# vignettes/start_here.Rmd:112) yields a "gms_lddecay" that predCrossVars hands to the device.
genmap2recombfreq_b200 <- function(m, nChr = NULL) structure(list(cM = m), class = "gms_recombfreq")
Ops.gms_recombfreq <- function(e1, e2) {
  if (.Generic == "*" && identical(e1, 2)) return(structure(e2, class = "gms_twoc"))
  stop("only 1-(2*c1) is supported on a lazy recombination-frequency matrix")
}
Ops.gms_twoc <- function(e1, e2) {
  if (.Generic == "-" && identical(e1, 1)) return(structure(e2, class = "gms_lddecay"))
  stop("only 1-(2*c1) is supported on a lazy recombination-frequency matrix")
}

.gms_chrom <- function(snps) {
  chr <- suppressWarnings(as.integer(sub("_.*$", "", snps)))      # "<chr>_<id>", R/helpers.R:199,208-211
  if (anyNA(chr) || !all(grepl("_", snps))) chr <- rep(0L, length(snps))
  chr
}

# Device-resident session: tiles + haplotypes of `parents`; reuse it across calls (CV folds).
# `devices`: integer vector of CUDA devices; with more than one, every device gets the tiles + haplotypes
# and predCrossVars() shards the cross list over them (gms_predict_sharded), all from this one R process.
# `map_scan = TRUE` switches on the map-structured fast path when recombFreqMat is a lazy map object.
gms_session <- function(haploMat, recombFreqMat, parents = NULL, devices = 0L, map_scan = FALSE) {
  if (length(devices) > 1) {
    ss <- lapply(devices, function(d) gms_session(haploMat, recombFreqMat, parents, d, map_scan))
    return(structure(list(handle = ss[[1]]$handle, handles = lapply(ss, `[[`, "handle"),
                          snps = ss[[1]]$snps, parents = ss[[1]]$parents), class = "gms_session"))
  }
  device <- devices
  snps <- colnames(haploMat)
  chr <- .gms_chrom(snps)
  ord <- order(match(chr, unique(chr)))                          # group by chromosome, stable
  snps <- snps[ord]; chr <- chr[ord]
  m_c <- as.integer(table(factor(chr, levels = unique(chr))))
  h <- .Call("gmsr_create", as.integer(device))
  if (inherits(recombFreqMat, "gms_lddecay")) {
    .Call("gmsr_set_genmap", h, m_c, as.double(recombFreqMat$cM[snps]))
  } else {
    ends <- cumsum(m_c); starts <- ends - m_c + 1L
    R <- recombFreqMat[snps, snps, drop = FALSE]                  # align BY NAME (R/predCrossVar.R:115)
    offzero <- all(vapply(seq_along(m_c), function(c) {
      i <- starts[c]:ends[c]; all(R[i, -i] == 0) }, logical(1)))
    blocks <- if (offzero) lapply(seq_along(m_c), function(c) { i <- starts[c]:ends[c]; R[i, i, drop = FALSE] })
              else list(R)                                        # general matrix: one block
    .Call("gmsr_set_recomb_blocks", h, lapply(blocks, function(b) { storage.mode(b) <- "double"; b }))
  }
  if (is.null(parents)) parents <- unique(sub("_Hap[AB]$", "", rownames(haploMat)))
  rows <- as.vector(rbind(paste0(parents, "_HapA"), paste0(parents, "_HapB")))
  H <- haploMat[rows, snps, drop = FALSE]                         # errors if a parent is missing, like :46
  storage.mode(H) <- "integer"
  .Call("gmsr_set_haplotypes", h, H)
  if (map_scan && inherits(recombFreqMat, "gms_lddecay")) .Call("gmsr_set_mode", h, 1L)
  structure(list(handle = h, handles = list(h), snps = snps, parents = as.character(parents)), class = "gms_session")
}

.gms_effmat <- function(EffectList, snps) {
  # n x m matrices with colnames = SNP ids. predType = "VPM" uses the posterior MEAN effects: the reference centres
  # each matrix and keeps the "scaled:center" attribute (R/predCrossVar.R:32-43), i.e. the column means - the row itself
  # for the usual 1 x m matrix, the mean over posterior samples for an n x m one
  do.call(cbind, lapply(EffectList, function(e) as.double(colMeans(e[, snps, drop = FALSE]))))
}

.gms_run <- function(session, CrossesToPredict, model, AddEffectList, DomEffectList = NULL, AlleleSubEffectList = NULL) {
  add <- .gms_effmat(AddEffectList, session$snps)
  dom <- if (model >= 1L) .gms_effmat(DomEffectList, session$snps) else NULL
  asub <- if (model == 2L) .gms_effmat(AlleleSubEffectList, session$snps) else NULL
  sire <- match(as.character(CrossesToPredict$sireID), session$parents) - 1L
  dam <- match(as.character(CrossesToPredict$damID), session$parents) - 1L
  if (anyNA(sire) || anyNA(dam)) stop("subscript out of bounds: parent not in haploMat")
  if (length(session$handles) > 1)
    return(.Call("gmsr_predict_sharded", session$handles, model, add, dom, asub, sire, dam))
  .Call("gmsr_predict", session$handle, model, add, dom, asub, sire, dam)
}

predCrossVars <- function(CrossesToPredict, modelType, AddEffectList, DomEffectList = NULL,
                          predType = "VPM", haploMat, recombFreqMat, ncores = 1, nBLASthreads = NULL,
                          session = NULL, ...) {
  starttime <- proc.time()[3]
  if (predType != "VPM") stop("predType='PMV' is not supported")
  model <- if (modelType == "AD") 1L else 0L                      # anything else behaves as "A" (:33,39,119,216)
  if (is.null(session)) {
    parents <- union(CrossesToPredict$sireID, CrossesToPredict$damID)          # :44-46
    session <- gms_session(haploMat, recombFreqMat, parents)
    on.exit(lapply(session$handles, function(h) .Call("gmsr_close", h)))
  }
  res <- .gms_run(session, CrossesToPredict, model, AddEffectList, DomEffectList)
  traits <- names(AddEffectList)
  pairs <- rbind(cbind(traits, traits), if (length(traits) > 1) t(combn(traits, 2)))   # :128-138
  comps <- c("VarA", "VarD")[seq_len(model + 1L)]
  totcomputetime <- proc.time()[3] - starttime
  keep <- which(res$nseg > 0)                                     # zero-segregation crosses vanish (:156-160,70-71)
  out <- tibble::as_tibble(CrossesToPredict)[keep, ]
  out$Nsegsnps <- res$nseg[keep]
  out$ComputeTime <- totcomputetime / max(1, nrow(CrossesToPredict))
  out$predVars <- lapply(keep, function(x) tibble::tibble(
    Trait1 = rep(pairs[, 1], each = length(comps)), Trait2 = rep(pairs[, 2], each = length(comps)),
    predOf = rep(comps, times = nrow(pairs)), predVar = as.vector(res$out[, , x])))
  print(paste0("Done predicting fam vars. Took ", round(totcomputetime / 60, 2), " mins for ", nrow(out), " crosses"))
  out
}
