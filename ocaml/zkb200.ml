(* OCaml side of the drop-in (SOURCE ONLY: not compiled in the build image, no OCaml there).

   Keeps the reference's functor signatures: [Curve.S], [Protocol.S],
   [Groth16.Make (C : Curve.S)], [Pinocchio.Make (C)] — only the bodies of the hot functions
   change.  The aliases at the bottom make the README's spelling ([Ecp.Bls12_381],
   [Protocol.Test]) resolve as well as the real one ([Curve.Bls12_381], [Test.Make]). *)

external init : int -> unit = "zkb200_init"
external g1_msm : bytes -> bytes -> bytes = "zkb200_g1_msm"
external g2_msm : bytes -> bytes -> bytes = "zkb200_g2_msm"
external qap_load : bytes -> bytes -> bytes -> bytes -> int -> int -> int64
  = "zkb200_qap_load_bytecode" "zkb200_qap_load_native"
external groth16_prove : int64 -> int64 -> bytes -> bytes -> bytes -> bytes = "zkb200_groth16_prove"
external pinocchio_prove : int64 -> int64 -> bytes -> bytes -> bytes = "zkb200_pinocchio_prove"

open Zukelang

(* [Curve.Bls12_381] with the MSM-shaped members of ExtendMap (curve.ml:79-119) rerouted. *)
module Bls12_381 = struct
  include Curve.Bls12_381

  let fr_bytes (xs : Fr.t list) = Bytes.concat Bytes.empty (List.map Fr.to_bytes xs)

  module G1 = struct
    include Curve.Bls12_381.G1

    let msm (pts : t list) (ks : Fr.t list) : t =
      let bases = Bytes.concat Bytes.empty (List.map to_bytes pts) in
      of_bytes_exn (Bytes.sub (g1_msm bases (fr_bytes ks)) 0 96)

    (* curve.ml:94-103 *)
    let dot m c =
      if not (Var.Set.equal (Var.Map.domain m) (Var.Map.domain c)) then begin
        prerr_endline "Domain mismatch"; assert false end;
      let ks = List.map (fun (k, _) -> Var.Infix.(c #! k)) (Var.Map.bindings m) in
      msm (List.map snd (Var.Map.bindings m)) ks

    (* curve.ml:112-118 *)
    let apply_powers (cs : Fr.t Polynomial.t) xis =
      if List.length cs > List.length xis then invalid_arg "apply_powers";
      msm (Misclib.List.take (List.length cs) xis) cs
  end
  (* G2 is the same with g2_msm / 192-byte points. *)
end

module Ecp = struct module Bls12_381 = Bls12_381 end   (* README.md:36-40 spelling *)
